// helper_cuda.h shim: the two things the reference's callers use (helper_cuda.h:999-1014,1244-1280):
// checkCudaErrors(call) and findCudaDevice(argc, argv) honouring "device=<n>" / "-device=<n>".
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define checkCudaErrors(val) cudamat_compat_check((val), #val, __FILE__, __LINE__)
inline void cudamat_compat_check(cudaError_t result, const char *func, const char *file, int line) {
    if (result != cudaSuccess) {
        fprintf(stderr, "CUDA error at %s:%d code=%d(%s) \"%s\" \n", file, line, (int)result, cudaGetErrorName(result), func);
        cudaDeviceReset();
        exit(EXIT_FAILURE);
    }
}

inline int findCudaDevice(int argc, const char **argv) {
    int dev = -1, count = 0;
    for (int k = 1; k < argc; ++k) {
        const char *a = argv[k];
        while (*a == '-') ++a;
        if (strncmp(a, "device=", 7) == 0) dev = atoi(a + 7);
    }
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        fprintf(stderr, "CUDA error: no devices supporting CUDA.\n");
        exit(EXIT_FAILURE);
    }
    if (dev < 0) {
        // pick the device with the most multiprocessors (stand-in for gpuGetMaxGflopsDeviceId)
        int best = -1;
        for (int k = 0; k < count; ++k) {
            cudaDeviceProp p;
            if (cudaGetDeviceProperties(&p, k) == cudaSuccess && p.multiProcessorCount > best) { best = p.multiProcessorCount; dev = k; }
        }
    } else if (dev >= count) {
        fprintf(stderr, ">> findCudaDevice: device %d is not a valid GPU device <<\n", dev);
        exit(EXIT_FAILURE);
    }
    checkCudaErrors(cudaSetDevice(dev));
    cudaDeviceProp p;
    checkCudaErrors(cudaGetDeviceProperties(&p, dev));
    printf("GPU Device %d: \"%s\" with compute capability %d.%d\n\n", dev, p.name, p.major, p.minor);
    return dev;
}
