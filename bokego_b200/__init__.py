"""bokego_b200 -- B200 (sm_100a) implementation of BokeGo's batched leaf-evaluation / playout hot path.

Host surface: `bokego_b200.go` and `bokego_b200.nnet` mirror the reference's `bokego.go` / `bokego.nnet`
modules; `bokego_b200.batched` holds the batched entry points that carry the throughput.  All compute
goes through hand-written CUDA kernels behind the C ABI in include/bokego_b200.h; there is no CPU path.
"""
__version__ = "0.1.0"
