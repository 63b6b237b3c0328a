"""smallz4_b200 -- B200-native optimal-parse LZ4 compressor (drop-in for smalLZ4's smallz4::lz4)."""
__version__ = "0.1.0"
