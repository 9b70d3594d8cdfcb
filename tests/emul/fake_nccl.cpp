// fake_nccl.cpp — TEST INFRASTRUCTURE: the six NCCL entry points libmdgpu binds at run time (dlopen) for its multi-device exchange step,
// implemented on host memory for the emulated library (tests/emul): "devices" are heap allocations of one process, so a reduce is an
// element-wise sum of the ranks' send buffers into the root's receive buffer, carried out when the outermost group ends. Loaded through
// $MDGPU_NCCL_LIB by tests/test_emulated_library.py only; the product never sees it.
// Built with -DMDG_LOOPBACK_CUDA (build_emul.build_loopback_nccl) the buffers are real device memory: the ranks' send buffers are staged through
// the host with cudaMemcpy. That variant lets tests/test_gpu_parity.py run the multi-device plan with both "devices" on ONE physical GPU
// (MDGPU_ALLOW_DUPLICATE_DEVICES=1), which real NCCL refuses — it exercises the library's peer plans, per-device threads, device switching and
// the exchange bookkeeping on hardware, not NCCL itself.
#ifdef MDG_LOOPBACK_CUDA
#include <cuda_runtime.h>
#endif
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <vector>

struct Call { const void* send; void* recv; size_t count; int type; int root; };
struct Comm { int rank, nranks; std::vector<Call> calls; std::vector<Comm*>* all; };
static int g_depth = 0; static std::vector<std::vector<Comm*>*> g_groups;

template <typename T> static void sum_into(const std::vector<Comm*>& all, size_t k) {
    const Call& r0 = all[0]->calls[k]; const int root = r0.root;
    std::vector<T> acc(r0.count, T(0));
#ifdef MDG_LOOPBACK_CUDA
    std::vector<T> tmp(r0.count);
    for (Comm* c : all) { cudaMemcpy(tmp.data(), c->calls[k].send, sizeof(T) * r0.count, cudaMemcpyDefault); for (size_t i = 0; i < r0.count; ++i) acc[i] += tmp[i]; }
    cudaMemcpy(all[(size_t)root]->calls[k].recv, acc.data(), sizeof(T) * r0.count, cudaMemcpyDefault);
#else
    for (Comm* c : all) { const T* s = (const T*)c->calls[k].send; for (size_t i = 0; i < r0.count; ++i) acc[i] += s[i]; }
    memcpy(all[(size_t)root]->calls[k].recv, acc.data(), sizeof(T) * r0.count);
#endif
}

extern "C" {
const char* ncclGetErrorString(int) { return "fake nccl"; }
int ncclCommInitAll(Comm** comms, int n, const int*) {
    auto* all = new std::vector<Comm*>(); g_groups.push_back(all);
    for (int r = 0; r < n; ++r) { comms[r] = new Comm{ r, n, {}, all }; all->push_back(comms[r]); }
    return 0;
}
int ncclCommDestroy(Comm* c) {
    for (auto it = c->all->begin(); it != c->all->end(); ++it) if (*it == c) { c->all->erase(it); break; }
    delete c; return 0;
}
int ncclGroupStart() { ++g_depth; return 0; }
int ncclReduce(const void* send, void* recv, size_t count, int type, int /*op: sum*/, int root, Comm* comm, void* /*stream*/) {
    comm->calls.push_back(Call{ send, recv, count, type, root });
    return 0;
}
int ncclGroupEnd() {
    if (--g_depth > 0) return 0;
#ifdef MDG_LOOPBACK_CUDA
    cudaDeviceSynchronize();   // the collectives were "enqueued" on stream 0 of each rank's device: everything before them must have finished
#endif
    for (auto* all : g_groups) {
        if (all->empty()) continue;
        const size_t n = (*all)[0]->calls.size();
        for (Comm* c : *all) if (c->calls.size() != n) return 1;   // every rank must issue the same collectives
        for (size_t k = 0; k < n; ++k) {
            switch ((*all)[0]->calls[k].type) {
                case 3: sum_into<uint32_t>(*all, k); break;   // ncclUint32
                case 5: sum_into<uint64_t>(*all, k); break;   // ncclUint64
                case 7: sum_into<float>(*all, k); break;      // ncclFloat32
                default: return 2;
            }
        }
        for (Comm* c : *all) c->calls.clear();
    }
    return 0;
}
}
