"""TEST INFRASTRUCTURE ONLY.  Oracles for the rasteriser hot path:

  oracle.cpu   -- Oracle 2: multi-threaded C++ CPU restatement (oracle/gvv_oracle.cpp)
  oracle.ref   -- Oracle 1: the reference's own CUDA core compiled in place (oracle/_ref/libgvv_ref.so)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this
package; the product (gvv_differentiable_cuda_renderer_b200) never does.
"""
