// Microbenchmark: issue throughput of the pair-distance test with scalar FP32 vs packed f32x2 (FADD2/FMUL2/FFMA2, sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o pairtest_f32x2 pairtest_f32x2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b){ u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b){ u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template<int MODE> __global__ void __launch_bounds__(256) k(const float4* refs, int nref, int reps, const float* tx, const float* ty, const float* tz, float G00, float G11, float G22, float r2, unsigned* out){
  __shared__ float4 s_ref[256];
  s_ref[threadIdx.x] = refs[threadIdx.x];
  __syncthreads();
  int i = blockIdx.x*blockDim.x+threadIdx.x;
  unsigned cnt = 0;
  if (MODE == 0) {   // scalar, 2 targets per thread
    float x0=tx[2*i], x1=tx[2*i+1], y0=ty[2*i], y1=ty[2*i+1], z0=tz[2*i], z1=tz[2*i+1];
    for (int rep = 0; rep < reps; ++rep) {
      #pragma unroll 4
      for (int r = 0; r < nref; ++r) {
        float4 f = s_ref[r];
        float dx=__fsub_rn(f.x,x0), dy=__fsub_rn(f.y,y0), dz=__fsub_rn(f.z,z0);
        float d2=__fmaf_rn(G00,__fmul_rn(dx,dx),__fmaf_rn(G11,__fmul_rn(dy,dy),__fmul_rn(G22,__fmul_rn(dz,dz))));
        float ex=__fsub_rn(f.x,x1), ey=__fsub_rn(f.y,y1), ez=__fsub_rn(f.z,z1);
        float e2=__fmaf_rn(G00,__fmul_rn(ex,ex),__fmaf_rn(G11,__fmul_rn(ey,ey),__fmul_rn(G22,__fmul_rn(ez,ez))));
        cnt += (d2 <= r2) + (e2 <= r2);
      }
      x0 += 1e-7f;
    }
  } else {           // packed
    u64 X = pack(tx[2*i], tx[2*i+1]), Y = pack(ty[2*i], ty[2*i+1]), Z = pack(tz[2*i], tz[2*i+1]);
    u64 g00 = pack(G00,G00), g11 = pack(G11,G11), g22 = pack(G22,G22);
    for (int rep = 0; rep < reps; ++rep) {
      #pragma unroll 4
      for (int r = 0; r < nref; ++r) {
        float4 f = s_ref[r];
        u64 dx = sub2(pack(f.x,f.x), X), dy = sub2(pack(f.y,f.y), Y), dz = sub2(pack(f.z,f.z), Z);
        u64 d2 = fma2(g00, mul2(dx,dx), fma2(g11, mul2(dy,dy), mul2(g22, mul2(dz,dz))));
        float a,b; unpack(d2,a,b);
        cnt += (a <= r2) + (b <= r2);
      }
      float a,b; unpack(X,a,b); X = pack(a+1e-7f,b);
    }
  }
  out[i] = cnt;
}
int main(){
  const int nthreads = 148*8*256, nref = 256, reps = 64;
  std::vector<float4> h_ref(256); std::vector<float> h(3*2*nthreads);
  for (int i=0;i<256;++i) h_ref[i]=make_float4((i*37%256)/256.f,(i*91%256)/256.f,(i*53%256)/256.f,0);
  for (size_t i=0;i<h.size();++i) h[i]=((i*2654435761u)%100000)/100000.f;
  float4* d_ref; float* d_t; unsigned* d_out;
  cudaMalloc(&d_ref,256*16); cudaMalloc(&d_t,h.size()*4); cudaMalloc(&d_out,nthreads*4);
  cudaMemcpy(d_ref,h_ref.data(),256*16,cudaMemcpyHostToDevice); cudaMemcpy(d_t,h.data(),h.size()*4,cudaMemcpyHostToDevice);
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  unsigned chk[2];
  for (int mode=0; mode<2; ++mode) {
    float best=1e9;
    for (int it=0; it<5; ++it) {
      cudaEventRecord(a);
      if (mode==0) k<0><<<nthreads/256,256>>>(d_ref,nref,reps,d_t,d_t+2*nthreads,d_t+4*nthreads,9866.f,9866.f,9866.f,0.0101f,d_out);
      else         k<1><<<nthreads/256,256>>>(d_ref,nref,reps,d_t,d_t+2*nthreads,d_t+4*nthreads,9866.f,9866.f,9866.f,0.0101f,d_out);
      cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); if (ms<best) best=ms;
    }
    std::vector<unsigned> o(nthreads); cudaMemcpy(o.data(),d_out,nthreads*4,cudaMemcpyDeviceToHost);
    unsigned long long s=0; for (auto v:o) s+=v; chk[mode]=(unsigned)s;
    double tests = 2.0*nthreads*nref*reps;
    printf("{\"mode\": \"%s\", \"ms\": %.3f, \"pair_tests_per_s\": %.4g, \"checksum\": %llu}\n", mode?"f32x2":"scalar", best, tests/(best*1e-3), s);
  }
  printf("{\"equal\": %s}\n", chk[0]==chk[1]?"true":"false");
  return 0;
}
