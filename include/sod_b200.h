/* sod_b200.h — C ABI of libsod_b200.so: the B200 (sm_100a) data-parallel hot path of
 * lartpang/Distributed-SOD-Project.
 *
 * The reference is pure Python and has NO FFI layer of its own: its hot path is reached
 * through five Python call sites.  Each entry point below replaces the arithmetic behind one of them;
 * the citation on each says which (paths relative to the reference repository).  A maintainer of the
 * reference binds these with `ctypes` exactly as INTEGRATION.md shows.
 *
 * Conventions
 *  - plain C: raw device pointers, element counts as int64_t, dtype enums, `void* stream` = cudaStream_t.
 *  - the caller owns every buffer; the library allocates nothing and never synchronises the stream.
 *  - return value: 0 ok; <0 contract error (sod_strerror); >0 a cudaError_t from the launch.
 *  - re-entrant; callable from any host thread (PyTorch's autograd thread calls the *_bwd entries).
 *  - collective entries (anything taking a sod_comm) must be called in the same order on every rank.
 */
#ifndef SOD_B200_H
#define SOD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOD_ABI_VERSION 7
#define SOD_MAX_WORLD 8
#define SOD_MAX_SEGMENTS 16
#define SOD_COMM_MAX_BLOCKS 1024   /* flag rows per channel */
#define SOD_COMM_CHANNELS 4        /* barrier channels: 0 gradient reduce+SGD, 3 plain all-reduce (1,2 reserved; the
                                      SyncBN exchange uses tagged packets, not flags) */

typedef enum { SOD_F32 = 0, SOD_BF16 = 1, SOD_F16 = 2 } sod_dtype;

enum {
    SOD_OK = 0,
    SOD_EINVAL = -1,       /* null pointer / negative size / bad enum */
    SOD_EALIGN = -2,       /* pointer or offset not 16-byte aligned */
    SOD_EWORKSPACE = -3,   /* workspace too small */
    SOD_EUNSUPPORTED = -4, /* shape outside what the kernels cover (e.g. C % 8 != 0) */
    SOD_ECOMM = -5         /* bad communicator (world > SOD_MAX_WORLD, missing peer pointer …) */
};

int sod_version(void);
const char* sod_strerror(int code);
/* sm count / compute capability of the current device; the hot path refuses anything but sm_100 */
int sod_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * (3) BCE-with-logits + CEL, forward AND backward in one kernel.
 * Replaces: torch.nn.BCEWithLogitsLoss (train.py:203) + CEL.forward (loss/CEL.py:15-20) as summed by
 *           get_total_loss (utils/pipeline_ops.py:37-42), and their autograd backward (train.py:302).
 *   total = w_bce*BCE(reduction) + w_cel*CEL;  grad = grad_scale * d total / d logits
 * scalars_out[8] (device, fp32): bce, cel, total, sum_p, sum_t, sum_pt, bce_sum(unreduced), n
 * mode: 0 auto (resident when the tensor fits the grid's shared memory, ≈5.4 M elements, else ring-streamed), 1 force the
 *       shared-memory-resident single-read kernel, 2 force the ring-streamed kernel (producer warp + shared-memory ring;
 *       phase 2 re-streams what did not stay resident, newest first, with L2 eviction hints)
 * workspace: sod_loss_workspace_bytes() bytes, 16-byte aligned; contents need not be initialised.
 * ------------------------------------------------------------------------------------------------ */
size_t sod_loss_workspace_bytes(void);
int sod_loss_bce_cel_fwd_bwd(const void* logits, int logits_dtype, const void* mask, int mask_dtype,
                             void* grad_logits, int grad_dtype, float* scalars_out, int64_t n,
                             int reduction_sum, float w_bce, float w_cel, float grad_scale, float eps,
                             int mode, void* workspace, size_t workspace_bytes, void* stream);
/* grad *= *scale_dev (device scalar): only for the case where the loss is not the root of backward */
int sod_scale_by_device_scalar(void* grad, int dtype, int64_t n, const float* scale_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Symmetric-memory communicator (one process per GPU; all ranks on one NVSwitch domain).
 * Built by the host from torch.distributed._symmetric_memory (buffer_ptrs / multicast_ptr): `peer[r]`
 * is rank r's arena as mapped into THIS process, `mc` the multicast (NVLS) mapping of the same arena
 * or 0.  Replaces the NCCL communicator of dist.init_process_group (train.py:73-77) for the hot path.
 * The first sod_comm_flag_bytes() bytes of every arena are the signal area and must be zeroed once
 * (before the first collective, followed by a process-group barrier).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t rank;
    int32_t world;
    uint64_t peer[SOD_MAX_WORLD];
    uint64_t mc;
    uint64_t arena_bytes;
    uint32_t* error_flag;   /* device word in local memory; set non-zero on a barrier timeout */
    uint64_t timeout_cycles;/* bounded spin in SM clock cycles; 0 = default (~20 s for barriers, ~2 s for packets) */
    uint32_t* block_seq;    /* device array [SOD_COMM_CHANNELS*SOD_COMM_MAX_BLOCKS] in LOCAL memory, zeroed once:
                               per-block barrier sequence numbers, advanced by the kernels themselves, so the
                               collectives carry no host-side sequence number and replay inside CUDA graphs */
} sod_comm;
size_t sod_comm_flag_bytes(void);

/* ------------------------------------------------------------------------------------------------
 * (1) gradient all-reduce ⊕ unscale ⊕ SGD-momentum.
 * Replaces: apex DDP(delay_allreduce=True) flat all-reduce + ×1/W (train.py:185), amp unscale
 *           (train.py:299), torch.optim.SGD.step as configured by make_optimizer
 *           (utils/pipeline_ops.py:295-313; train.py:303), and optimizer.zero_grad (train.py:297).
 * Flat fp32 buffers of n elements (n % 4 == 0), laid out group-major; `segs` (host array) gives each
 * contiguous range its lr / weight_decay / momentum; SOD_SEG_FROZEN ranges receive the averaged
 * gradient but p and v are left alone (the reference leaves `div_2.*` out of every param group).
 *   g = (Σ_ranks grad) * inv_scale / world + wd*p ;  v = mu*v + g ;  p -= lr*v
 * world==1 (comm NULL): purely local.  world>1: grad and param live in the arena at grad_off/param_off;
 * rank r reduces shard r (NVLS multimem.ld_reduce when comm->mc != 0 and algo allows, else peer loads
 * in rank order), updates its shard of p and v, and writes the new p to every rank (multimem.st / peer
 * stores).  `mom` is the local momentum buffer, full length n (only shard r is used when world>1).
 * found_inf (device, may be NULL): when non-NULL and *found_inf != 0 the whole step is skipped.
 * flags: SOD_SGD_ZERO_GRAD zeroes the local gradient buffer(s) on the way out.
 * Mixed precision ("fp32 master" of apex amp, train.py:183): `grad16` (bf16, same flat layout, may be NULL) holds
 * the gradients autograd produced in bf16 — the kernel adds them to `grad` in registers, so no bf16→fp32 cast
 * kernel and no fp32 accumulation kernel run per tensor; `shadow16` (bf16, may be NULL) receives the rounded copy
 * of every updated parameter, which the next forward's convolutions consume directly (no per-iteration weight
 * cast).  Ranges flagged SOD_SEG_GRAD16 take their gradient from `grad16` alone.  With world>1 `grad16` lives in the
 * arena too (grad16_off) and those ranges are reduced from the bf16 values themselves; sod_grad_merge_bf16 is only needed
 * when fp32 gradients were ALSO written for such a range (a forward outside autocast).  `shadow16` is the LOCAL bf16
 * buffer every rank refreshes from the all-gathered parameters.
 * ------------------------------------------------------------------------------------------------ */
enum { SOD_SEG_FROZEN = 1,
       SOD_SEG_GRAD16 = 2 /* the gradient of this range exists ONLY in the bf16 buffer (`grad16`): the fp32 buffer is neither read
                             nor cleared there, and with world>1 the bf16 values themselves cross NVLink (summed in fp32) */ };
typedef struct {
    int64_t begin, end;        /* element range, multiples of 4 */
    float lr, weight_decay, momentum;
    int32_t flags;
} sod_sgd_segment;
enum { SOD_SGD_ZERO_GRAD = 1, SOD_ALGO_NO_MULTIMEM = 2,
       SOD_DEBUG_TIMING = 4 /* syncbn: per-CTA globaltimer stamps behind the workspace (tools/bn_phases.py) */,
       SOD_BN_ACCUMULATE_PARAM_GRADS = 8 /* syncbn_bwd: dgamma/dbeta += (write straight into the bound .grad) */,
       SOD_ALGO_FORCE_MULTIMEM = 16 /* use NVLS even at world 2, where the default is peer loads */,
       SOD_BN_BWD_MASK_FROM_X = 32 /* syncbn_bwd: re-derive the ReLU mask from x with the forward's arithmetic instead
                                      of reading y (one input stream less); needs relu, beta and dres == NULL */,
       SOD_BN_L2_HINTS = 64 /* syncbn_bwd: L2 eviction-priority hints on the bulk copies — evict-last for chunks that
                               are fetched twice, evict-first for last uses */,
       SOD_BN_LAUNCH_COOP = 128 /* syncbn_fwd/bwd: cooperative launch (driver-guaranteed co-residency of the grid) */,
       SOD_BN_LAUNCH_PDL = 256 /* syncbn_fwd/bwd: programmatic dependent launch (prologue overlaps the producer's tail) */ };

/* lr_dev (device, fp32[nseg], may be NULL): when given, the learning rate of segment i is read from lr_dev[i] at
 * execution time instead of segs[i].lr — a captured CUDA graph then follows CustomScheduler
 * (utils/pipeline_ops.py:225-229, stepped per epoch train.py:240-241 or per iteration train.py:288-289) with one
 * small H2D copy per change and no re-capture. */
int sod_sgd_momentum(float* param, float* mom, float* grad, void* grad16, void* shadow16, int64_t n,
                     const sod_sgd_segment* segs, int nseg, const float* lr_dev, float inv_scale,
                     const uint32_t* found_inf, int flags, void* stream);
/* grad16_off (world>1): arena offset of the symmetric bf16 gradient buffer, or 0.  Ranges flagged SOD_SEG_GRAD16 are reduced
 * from it (peer loads of the bf16 values, fp32 sum in rank order — half the reduce-scatter bytes and no merge pre-pass);
 * all other ranges from the fp32 buffer at grad_off as before. */
int sod_allreduce_sgd(const sod_comm* comm, uint64_t grad_off, uint64_t grad16_off, uint64_t param_off, float* mom,
                      void* shadow16, int64_t n, const sod_sgd_segment* segs, int nseg, const float* lr_dev,
                      float inv_scale, const uint32_t* found_inf, int flags, void* stream);
/* Multi-tensor gather of bf16 gradients into the flat bf16 gradient buffer (`grad16` above): item i copies
 * numel elements from the dense tensor `src` to dst16[dst_offset ...] (dst_offset % 8 == 0).  Replaces the
 * per-parameter accumulate kernels autograd launches at the end of backward (train.py:302): the weight gradients are
 * left where cuDNN wrote them and collected by ONE launch (per SOD_GATHER_MAX_ITEMS tensors). */
#define SOD_GATHER_MAX_ITEMS 160
typedef struct {
    const void* src;
    int64_t dst_offset, numel;
} sod_gather_item;
int sod_grad_gather16(const sod_gather_item* items, int nitems, void* dst16, int64_t dst_elems, void* stream);
/* grad[i] += float(grad16[i]); grad16[i] = 0   (world>1 pre-pass, one launch over the flat buffers) */
int sod_grad_merge_bf16(float* grad, void* grad16, int64_t n, void* stream);
/* *found_inf |= any(!isfinite(grad)) — the amp overflow check (train.py:299), one read of grad */
int sod_grad_nonfinite(const float* grad, int64_t n, uint32_t* found_inf, void* stream);

/* plain in-place SUM all-reduce of fp32 data at arena offset `off` (BASELINE config 5 sweep; also the
 * scalar loss mean of utils/tensor_ops.py:60-64 with scale = 1/W).  algo: 0 auto, 1 one-shot, 2 two-shot */
int sod_allreduce_f32(const sod_comm* comm, uint64_t off, int64_t n, float scale, int algo, int flags,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * (2) SyncBatchNorm, statistics exchange fused with normalize/affine (+ optional pre-add, residual, ReLU).
 * Replaces: apex convert_syncbn_model/SyncBatchNorm forward+backward (train.py:180) — local Welford +
 *           2 all_gathers + elementwise forward; reduce + 2 all_reduces + elementwise backward.
 * Data layout: channels-last matrices [M = N*H*W rows, C channels] (dtype ∈ bf16/f16/f32; C a power-of-two
 * multiple of 8, C ≤ 2048), fp32 gamma/beta/running stats/saved stats.
 *   z = x (+ pre_add);  mean/var over M*world rows;  y = relu?((z-mean)*invstd*gamma + beta (+ residual))
 * conv_bias1/2 (per-channel, dtype conv_bias_dtype, may be NULL): biases of the convolution(s) producing x / pre_add,
 *   folded in here — z = x (+ pre_add) + b1 + b2 — so that the bias-less convolution output needs no separate
 *   bias-add pass; the backward adds Σ_rows dz (the bias gradient, LOCAL like dgamma/dbeta) into dconv_bias1/2.
 * training=0: uses running stats, no exchange.  comm may be NULL (world 1).
 * stats_off: arena offset of the exchange slot for this call (sod_syncbn_exchange_bytes(C) bytes; the host
 * rotates ≥2 slots); seq/epoch: the packet tag — it must be unique among all syncbn calls (forward AND
 * backward) that share a workspace / slot and identical on all ranks for the same call.  epoch == NULL: tag =
 * seq (31 bits; one global host call counter does it).  epoch != NULL (CUDA-graph replay): tag =
 * 2^31 | (*epoch << 10) | (seq & 1023), with seq the call index inside the captured iteration and *epoch a
 * device counter the captured iteration increments once.
 * workspace: sod_syncbn_workspace_bytes(rows, C) bytes, zero-filled once, then owned by the library.
 * ------------------------------------------------------------------------------------------------ */
size_t sod_syncbn_workspace_bytes(int64_t rows, int channels);
size_t sod_syncbn_exchange_bytes(int channels);
int sod_syncbn_fwd(const void* x, const void* pre_add, const void* residual, void* y, int dtype,
                   const float* gamma, const float* beta, float* running_mean, float* running_var,
                   float* save_mean, float* save_invstd, int64_t rows, int channels, float momentum,
                   float eps, int relu, int training, const sod_comm* comm, uint64_t stats_off, uint32_t seq,
                   const uint32_t* epoch, int64_t* num_batches_tracked /* += 1 when training; may be NULL */,
                   const void* conv_bias1, const void* conv_bias2, int conv_bias_dtype, void* workspace,
                   size_t workspace_bytes, int flags, void* stream);
/* dz = d/d(x) = d/d(pre_add); dres = relu-masked dy (written only if non-NULL; may alias nothing);
 * dgamma/dbeta: LOCAL sums (the gradient all-reduce averages them with every other parameter).
 * y: the forward's output, read for the ReLU mask (may be NULL when relu == 0 or with SOD_BN_BWD_MASK_FROM_X);
 * beta: read only with SOD_BN_BWD_MASK_FROM_X (may be NULL otherwise). */
int sod_syncbn_bwd(const void* dy, const void* x, const void* pre_add, const void* y, void* dz, void* dres,
                   int dtype, const float* gamma, const float* beta, const float* save_mean,
                   const float* save_invstd, float* dgamma, float* dbeta, int64_t rows, int channels, int relu,
                   const sod_comm* comm, uint64_t stats_off, uint32_t seq, const uint32_t* epoch, const void* conv_bias1,
                   const void* conv_bias2, void* dconv_bias1, void* dconv_bias2, int conv_bias_dtype, void* workspace,
                   size_t workspace_bytes, int flags, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Resampling ops on either side of the SyncBN kernels, channels-last [N,H,W,C], C % 8 == 0.
 * sod_upsample2x_bilinear_*: bilinear ×2, align_corners=False — `cus_sample` / `upsample_add`
 *   (utils/tensor_ops.py:12-25; call sites network/TestModel.py:100-126, module/MyLightModule.py:42-52);
 *   `add` (may be NULL) is the tensor `upsample_add` sums with, already at the output size.  (h, w) = INPUT size.
 *   The backward is a gather (no atomics): deterministic.
 * sod_avgpool2x2_*: AvgPool2d((2,2), stride=2) — `h2l_pool` (module/MyLightModule.py:14); (h_out, w_out) = OUTPUT size.
 * ------------------------------------------------------------------------------------------------ */
int sod_upsample2x_bilinear_fwd(const void* x, const void* add, void* y, int n, int h, int w, int c, int dtype, void* stream);
int sod_upsample2x_bilinear_bwd(const void* dy, void* dx, int n, int h, int w, int c, int dtype, void* stream);
int sod_avgpool2x2_fwd(const void* x, void* y, int n, int h_out, int w_out, int c, int dtype, void* stream);
int sod_avgpool2x2_bwd(const void* dy, void* dx, int n, int h_out, int w_out, int c, int dtype, void* stream);

/* Column sum of a channels-last matrix [rows, c] (c a power-of-two multiple of 8, ≤ 2048) into out[c] (same dtype, fp32
 * accumulation, deterministic): the bias gradient Σ_rows dy of a convolution that does not feed a BatchNorm — the five
 * `trans*` 1x1 convolutions of network/TestModel.py:32-36 — which autograd otherwise computes with a generic reduction.
 * workspace: sod_colsum_workspace_bytes() bytes, zero-filled once, then owned by the library (one stream at a time). */
size_t sod_colsum_workspace_bytes(void);
int sod_colsum(const void* x, void* out, int64_t rows, int c, int dtype, void* workspace, size_t workspace_bytes, void* stream);

/* MaxPool2d(kernel 3, stride 2, padding 1) of the ResNet stem (backbone/origin/resnet.py `maxpool`; `div_4` in
 * backbone/origin/from_origin.py:7-15), channels-last [N,H,W,C], C % 8 == 0; (h, w) = INPUT size, output
 * ((h-1)/2+1, (w-1)/2+1).  argmax: one byte per OUTPUT element (window position kh*3+kw), written by the forward and
 * read by the backward, which is a deterministic gather with fp32 accumulation.  Ties / NaN as torch: first maximum
 * in (kh, kw) scan order, NaN propagates.  (Checked against torch on B200 in round 2 and enabled by default.) */
int sod_maxpool3x3s2_fwd(const void* x, void* y, void* argmax, int n, int h, int w, int c, int dtype, void* stream);
int sod_maxpool3x3s2_bwd(const void* dy, const void* argmax, void* dx, int n, int h, int w, int c, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Data-side neighbours of the iteration.
 *
 * sod_preprocess_batch — the per-sample tensor transforms and the multi-scale collate of the reference's loader as one
 * kernel over a uint8 batch: ToTensor + Normalize (utils/dataset.py:86-96), mask ToTensor (:110-116), optional
 * left-right flip (utils/joint_transforms.py:19-23; flip_u8[i] != 0 mirrors sample i), and the collate's
 * F.interpolate(img, bilinear, align_corners=False) / F.interpolate(mask, nearest) to (ho, wo)
 * (utils/dataset.py:125-132).  img_u8 [n,hs,ws,3] HWC, mask_u8 [n,hs,ws] (may be NULL together with out_mask),
 * out_img: storage [n,ho,wo,3] = a channels-last [n,3,ho,wo] tensor of out_dtype, out_mask fp32 [n,1,ho,wo].
 *
 * sod_saliency_* — sufficient statistics of the reference's evaluation metrics (utils/saliency_metric.py:8-239 as
 * driven by train.py:383-412): quantize = ToPILImage's mul(255).byte() (optionally after a sigmoid); head = per image
 * {min_u, max_u, max_gt, 0, n_fg, Σy, Σx, 0} (int64[8]); hist = per image uint32[4 quadrants][2 gt][256 k] joint
 * histogram of k = u - min_u, quadrants split at split_yx[2i], split_yx[2i+1] (rows < y / cols < x first), counters
 * ADDED to (zero them first).  Integer arithmetic only: results are exact and order independent.
 * ------------------------------------------------------------------------------------------------ */
int sod_preprocess_batch(const void* img_u8, const void* mask_u8, const void* flip_u8, void* out_img, int out_dtype,
                         float* out_mask, int n, int hs, int ws, int ho, int wo, const float* mean3,
                         const float* std3, void* stream);
int sod_saliency_quantize(const void* pred, int dtype, void* out_u8, int64_t n, int apply_sigmoid, void* stream);
int sod_saliency_head(const void* pred_u8, const void* gt_u8, int n, int h, int w, int64_t* head, void* stream);
int sod_saliency_hist(const void* pred_u8, const void* gt_u8, int n, int h, int w, const int64_t* head,
                      const int32_t* split_yx, uint32_t* hist, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SOD_B200_H */
