// fake_cudart.cpp — TEST INFRASTRUCTURE: the two dozen CUDA runtime entry points the product's host code calls, implemented on host memory
// with immediate execution, so that the WHOLE library (plan.cu included) can be linked against the emulated kernels and driven through its
// C ABI in the CPU test stage (tests/test_emulated_library.py). "Device" memory is heap memory; a stream runs its work at the call; events
// carry no time. This object is linked into tests/emul/build/libmdgpu_emul.so only — the product never sees it and has no CPU path.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

extern "C" {
cudaError_t cudaGetDeviceCount(int* n) { const char* e = getenv("MDGPU_EMUL_DEVICES"); *n = e ? atoi(e) : 1; if (*n < 1) *n = 1; return cudaSuccess; }   // several "devices" share the heap
cudaError_t cudaDeviceGetPCIBusId(char*, int, int) { return cudaErrorNotSupported; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) { *v = (a == cudaDevAttrMultiProcessorCount) ? 4 : 0; return cudaSuccess; }   // 4 "SMs": default batch of 4 frames
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "emulated runtime"; }
cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { memset(a, 0, sizeof(*a)); a->type = cudaMemoryTypeUnregistered; return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; ++r) memmove((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)calloc(1, 8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free((void*)s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)calloc(1, 8); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)calloc(1, 8); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free((void*)e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
}
