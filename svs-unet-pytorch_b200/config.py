"""Constants of the separation path (mirror of reference config.py:47-51, the active "1209" set).

The CUDA kernels are specialised for exactly these values; the older parameter sets kept as
comments in the reference (config.py:11-44: hop 256, 44.1 kHz, INPUT_LEN 512/1536) are rejected
with an error by the library rather than served by a fallback."""


def num2str(n):
    """reference config.py:1-9 (zero-pad to 4 digits)."""
    return str(n).zfill(4)


WINDOW_SIZE = 1024
HOP_SIZE = 768
SAMPLE_RATE = 8192
INPUT_LEN = 128
SAMPLES_PER_SONG = 64

N_BINS = WINDOW_SIZE // 2 + 1          # 513 rows of a *_spec.npy
PATCH_BINS = 512                       # DC row dropped before the net (reference inference.py:68)
