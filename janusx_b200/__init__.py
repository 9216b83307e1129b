"""janusx_b200 -- B200-native (sm_100a) exact LMM association scan behind the JanusX GWAS API.

Only the hot path of `jx gwas -lmm / -lmm2 / -fvlmm` lives here:
  csrc/      hand-written CUDA kernels + the C ABI (include/jxb200.h) -> libjxb200.so
  jxrs       the `janusx.janusx` function surface for this path (ctypes over the C ABI)
  assoc      LMM / LMM2 / FvLMM model objects (python/janusx/pyBLUP/assoc.py interface)
  dist       SNP-range sharding across GPUs (torch.distributed / NCCL for the one broadcast)
  synth      synthetic inputs with the reference simulator's distributions
There is no CPU implementation in this package: compute calls raise without libjxb200.so + a GPU.
"""
__version__ = "0.1.0"
