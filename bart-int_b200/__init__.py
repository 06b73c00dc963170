"""bartint_b200: B200-native (sm_100a) implementation of the BART-Int posterior-evaluation hot path.

Layout
  csrc/            hand-written CUDA kernels + the C ABI (include/bartint.h) -> libbartint_cuda.so
  _lib.py          ctypes binding of the C ABI (fails loudly when the library is missing)
  forest.py        Forest / Points handles over the ABI, FlatForest (SoA) container
  api.py           host-side mirror of the reference's R interface for this path
                   (sampleIntegrals, predict, BARTInt, computeBART ... same names and semantics)
  dist.py          draw-/point-sharded multi-GPU evaluation (one process per GPU, torch.distributed)
  synth.py         synthetic posteriors and Genz-shaped data for tests and the benchmark
  R/               the .Call shim sources a BART-Int maintainer adds (not buildable here: no R)
"""
from .forest import (FlatForest, Forest, Points, StagedMatrix, WireStrings, BartIntError,  # noqa: F401
                     design_step_dbarts, init_group, group_size, group_exchange, shutdown_group, topk)
from . import api, synth  # noqa: F401

__all__ = ["FlatForest", "Forest", "Points", "StagedMatrix", "WireStrings", "BartIntError", "design_step_dbarts", "init_group",
           "group_size", "group_exchange", "shutdown_group", "topk", "api", "synth"]
