"""B200-native data-parallel hot path for lartpang/Distributed-SOD-Project."""
__version__ = "0.1.0"
