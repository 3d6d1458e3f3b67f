"""srsran_4g_b200 - host-side Python mirror of the B200-native LTE turbo-decode engine.

The product is the C-ABI shared library libsrsran_b200.so (include/srsran_b200.h); this package only loads it through
ctypes and mirrors the reference's operator interface for this path (srsran_tdec_*, srsran_rm_turbo_rx_lut, decode_tb)
so that the parity tests read like the reference's own tests. There is no CPU fallback: without the built library or
without a CUDA device every compute call raises."""
from .binding import (CRC_24A, CRC_24B, CRC_NONE, Engine, Multi, SrsB200Error, Tdec, TransportBlock, cbindex, cbsegm, cbsize, lib, lib_path,
                      rm_table, tdec8_windows)

__all__ = ["Engine", "Multi", "Tdec", "TransportBlock", "SrsB200Error", "CRC_NONE", "CRC_24A", "CRC_24B", "cbsize", "cbindex", "cbsegm",
           "rm_table", "lib", "lib_path", "tdec8_windows"]
