// die_probe.cu -- which L2 die is each SM on?  B200 is two dies; an L2 hit costs ~234 cycles from the near die and ~262 from the
// far one (B300_MICROARCH.md), and addresses are homed on one die at 2 KB grain.  Every SM times dependent L2-hitting loads to ONE
// 2 KB region: the latencies fall into two groups.  Also prints which SMs the CTAs of a cluster-of-2 launch land on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/die_probe tools/die_probe.cu && tools/bin/die_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

__global__ void probe(const unsigned *buf, unsigned *lat, unsigned *smid_of_block, int n_regions, int region_stride_words) {
    extern __shared__ char pad[];
    if (threadIdx.x != 0) return;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    smid_of_block[blockIdx.x] = smid;
    for (int r = 0; r < n_regions; r++) {
        const unsigned *p = buf + (size_t)r * region_stride_words;
        unsigned idx = 0, best = 1u << 30, sink = 0;
        for (int rep = 0; rep < 64; rep++) {
            unsigned long long t0, t1;
            unsigned v;
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) : "r"(idx) : "memory");
            asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p + idx) : "memory");   // .cg: L2 only
            asm volatile("and.b32 %0, %1, 255;" : "=r"(idx) : "r"(v) : "memory");  // issues only when the load has landed (in-order issue)
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) : "r"(idx) : "memory");
            unsigned d = (unsigned)(t1 - t0);
            if (rep > 8 && d < best) best = d;
            sink += idx;  // keeps the load chain alive in ptxas
        }
        lat[smid * n_regions + r] = best;
        if (sink == 0xFFFFFFFFu) smid_of_block[blockIdx.x] = sink;  // never true (the buffer holds zeros)
    }
}

int main() {
    int n_sm = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    const int n_regions = 8, stride_words = (1 << 20) / 4;  // 8 regions 1 MB apart
    unsigned *buf, *lat, *smid_of_block;
    cudaMalloc(&buf, (size_t)n_regions * stride_words * 4);
    cudaMemset(buf, 0, (size_t)n_regions * stride_words * 4);
    cudaMallocManaged(&lat, n_sm * n_regions * 4);
    cudaMallocManaged(&smid_of_block, n_sm * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(n_sm);
    cfg.blockDim = dim3(32);
    cfg.dynamicSmemBytes = 200 * 1024;  // one CTA per SM
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; rep++) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, probe, (const unsigned *)buf, lat, smid_of_block, n_regions, stride_words);
        cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return 1; }
    }
    printf("latency (cycles) per smid for %d regions:\n", n_regions);
    for (int s = 0; s < n_sm; s++) {
        printf("smid %3d:", s);
        for (int r = 0; r < n_regions; r++) printf(" %4u", lat[s * n_regions + r]);
        printf("\n");
    }
    // die by region 0: split at the midpoint of min and max
    unsigned lo = 1u << 30, hi = 0;
    for (int s = 0; s < n_sm; s++) { lo = std::min(lo, lat[s * n_regions]); hi = std::max(hi, lat[s * n_regions]); }
    unsigned mid = (lo + hi) / 2;
    printf("region 0: min %u max %u -> split at %u\n", lo, hi, mid);
    printf("cluster (pair) -> smids, die of each (by region 0):\n");
    int same = 0, changes = 0, prev = -1;
    for (int b = 0; b < n_sm; b += 2) {
        int d0 = lat[smid_of_block[b] * n_regions] > mid, d1 = lat[smid_of_block[b + 1] * n_regions] > mid;
        printf("pair %2d: smid %3u,%3u die %d,%d\n", b / 2, smid_of_block[b], smid_of_block[b + 1], d0, d1);
        same += d0 == d1;
        if (prev >= 0 && prev != d0) changes++;
        prev = d0;
    }
    printf("pairs on one die: %d of %d; die changes between consecutive pairs: %d\n", same, n_sm / 2, changes);
    return 0;
}
