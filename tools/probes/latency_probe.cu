// Latency of one dependent global load as seen by a lightly loaded GPU, as a function of the footprint the
// addresses are spread over (TLB reach) -- the regime of the ordered resolution kernels (few CTAs, cold data).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency_probe latency_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(const double *buf, size_t n, int hops, unsigned long long *out, double *sink, unsigned seed)
{
    unsigned long long r = (blockIdx.x * 2654435761u + seed) | 1u;
    double acc = 0;
    long long t0 = clock64();
    for (int h = 0; h < hops; h++) {
        r = r * 6364136223846793005ull + 1442695040888963407ull;
        size_t i = (size_t)((r >> 20) % n);
        acc += buf[i];
        r += (unsigned long long)(acc != 12345.0); // make the next address depend on the load
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { atomicAdd(out, (unsigned long long)(t1 - t0)); sink[blockIdx.x] = acc; }
}
__global__ void touch(double *buf, size_t n) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (i < n) buf[i] = 1.0; }
int main()
{
    size_t nmax = (size_t)1 << 30; // 8 GB of doubles
    double *buf, *sink; unsigned long long *out;
    cudaMalloc(&buf, nmax * 8); cudaMalloc(&sink, 4096 * 8); cudaMalloc(&out, 8);
    touch<<<(unsigned)((nmax + 255) / 256), 256>>>(buf, nmax);
    cudaDeviceSynchronize();
    for (size_t mb : {16, 64, 256, 1024, 2048, 4096, 8192}) {
        size_t n = mb * 1024 * 1024 / 8;
        for (int ctas : {148, 720}) {
            const int hops = 8;
            for (int rep = 0; rep < 3; rep++) {
                cudaMemset(out, 0, 8);
                touch<<<(unsigned)(((size_t)1 << 27) / 256), 256>>>(buf + (nmax - ((size_t)1 << 27)), (size_t)1 << 27); // churn L2 (1 GB)
                probe<<<ctas, 32>>>(buf, n, hops, out, sink, 17 + rep);
                cudaDeviceSynchronize();
                unsigned long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
                if (rep == 2) printf("footprint %5zu MB  ctas %4d  cycles per dependent load %.0f\n", mb, ctas, (double)c / ctas / hops);
            }
        }
    }
    return 0;
}
