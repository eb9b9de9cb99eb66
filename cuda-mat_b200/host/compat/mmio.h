// mmio.h shim: example.cpp:22 includes the NIST header but only calls loadMMSparseMatrix
// (mmio_wrapper.h). The loader itself lives in libcudamat_b200.so (csrc/mmload.cpp).
#pragma once
typedef char MM_typecode[4];
