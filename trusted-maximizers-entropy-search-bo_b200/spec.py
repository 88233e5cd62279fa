"""Constants of include/tes_spec.h mirrored for Python callers."""
STREAM_EP_Y = 1
STREAM_SP_Y = 2
STREAM_BSP_Y = 3
STREAM_BEP_Y = 4
