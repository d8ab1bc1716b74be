#!/bin/bash
# Memory-safety pass (compute-sanitizer is closed on the pool): builds the library with -DELX_DEBUG_BOUNDS (every shared-memory
# table access and slot index of the tiled kernels is range-checked on the device, a violation raises the handle's error flag)
# into /tmp and runs the free-running GPU tests against it (ELX_LIB selects the build).  Run under gpurun from the repo root.
set -e
SRC="elynx_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 -shared -cudart static -lz -DELX_DEBUG_BOUNDS \
  -o /tmp/libelynx_b200_dbg.so $SRC/elx_core.cu $SRC/elx_fast.cu $SRC/elx_pair.cu $SRC/elx_quad.cu $SRC/elx_duo.cu $SRC/elx_examine.cu \
  $SRC/elx_host.cpp $SRC/elx_model.cpp $SRC/elx_io.cpp
ELX_LIB=/tmp/libelynx_b200_dbg.so python -m pytest tests/test_gpu_duo.py tests/test_gpu_parity.py tests/test_gpu_configs.py -q -x -k "freerun or duo or cfg or quad or pair" 2>&1 | tail -4
