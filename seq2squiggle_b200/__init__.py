"""seq2squiggle_b200 — B200-native signal-generation path for ``seq2squiggle predict``.

Host side (Python/PyTorch for device memory and streams) mirrors the reference's predict interface:
``cli.main`` (``seq2squiggle predict``), ``inference.inference_run``, ``model.seq2squiggle`` and
``signal_io.BLOW5Writer``; the arithmetic runs in hand-written sm_100a CUDA kernels behind the C-ABI
declared in ``include/s2s_b200.h`` (``libs2s_b200.so``).  There is no CPU fallback.
"""
__version__ = "0.1.0"
