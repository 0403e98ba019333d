"""stm32f7-rtlsdr_b200 -- B200-native IQ sample-processing path behind the stm32f7-rtlsdr
buffer-callback boundary.

The product is ``libb200sdr.so`` (hand-written CUDA for sm_100a, plain C ABI declared in
``include/b200sdr.h``).  This package is only the thin ctypes host binding used by the tests
and by ``bench.py``; a C host driver links the library directly (see ``INTEGRATION.md``).

There is NO CPU fallback: importing works without a GPU (so the symbol table can be checked),
but every compute entry point needs a CUDA device and the built library, and fails loudly
otherwise.
"""
from .binding import (  # noqa: F401
    B200Sdr,
    B200SdrError,
    Config,
    LIB_PATH,
    CHAIN_SPECTRUM,
    CHAIN_WBFM,
    CHAIN_AM,
    CHAIN_COUNTER,
    WINDOW_RECT,
    WINDOW_HANN,
    WINDOW_BLACKMAN,
    AVG_MEAN,
    AVG_EMA,
    SYNTH_COUNTER,
    SYNTH_MULTITONE,
    SYNTH_WBFM,
    SYNTH_AM,
    FIR_ENGINE_FP32,
    FIR_ENGINE_TENSOR,
    OK,
    BUSY,
    FAIL,
    NOT_SUPPORTED,
    load_library,
    declared_symbols,
    spectrum_frames,
    wbfm_disc_len,
    wbfm_audio_len,
    wbfm_stream_audio_len,
    stream_chunk_samples,
    am_audio_len,
    synth_fill_host,
)
