// write_ceiling.cu -- the write-only HBM bandwidth this B200 reaches, the ceiling the observation-emitting kernels are held
// against.  Three writers over 0.4 .. 3.2 GB: cudaMemsetAsync, a grid-stride kernel of 16-byte stores,
// and a kernel that stages 16 KB tiles in shared memory and emits them with 1-D TMA bulk stores
// (cp.async.bulk.global.shared::cta), the store path of the fused step + observation kernel.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/write_ceiling tools/write_ceiling.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void fill16(uint4 *p, size_t n16) {
    uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

template <int TILE>
__global__ void fill_bulk(unsigned char *p, size_t ntiles) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < TILE / 16; i += blockDim.x) ((uint4 *)sm)[i] = make_uint4(1, 2, 3, 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t src = (uint32_t)__cvta_generic_to_shared(sm);
        int pending = 0;
        for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + t * TILE), "r"(src), "r"(TILE) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++pending == 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); pending = 4; }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// the pattern of the fused step + observation kernel: every CTA's thread 0 emits its tiles as `per_item` consecutive TILE-byte bulk
// stores per work item (a game's block of zeros), items handed out round-robin
template <int TILE>
__global__ void fill_bulk_items(unsigned char *p, size_t nitems, int per_item) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < TILE / 16; i += blockDim.x) ((uint4 *)sm)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        uint32_t src = (uint32_t)__cvta_generic_to_shared(sm);
        int warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
        for (size_t it = (size_t)blockIdx.x * wpc + warp; it < nitems; it += (size_t)gridDim.x * wpc) {
            unsigned char *d = p + it * (size_t)per_item * TILE;
            for (int k = 0; k < per_item; k++)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + (size_t)k * TILE), "r"(src), "r"(TILE) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // as the kernel does before it scatters values over the zeros
        }
    }
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const size_t sizes[] = {400u << 20, 800u << 20, 1600u << 20, 3200ull << 20};
    unsigned char *buf; cudaMalloc(&buf, sizes[3]);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaFuncSetAttribute(fill_bulk<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    cudaFuncSetAttribute(fill_bulk<65536>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"unit\": \"GB/s\", \"rows\": [\n", prop.name, prop.multiProcessorCount);
    for (int si = 0; si < 4; si++) {
        size_t n = sizes[si];
        float best[5] = {0, 0, 0, 0, 0};
        for (int rep = 0; rep < 12; rep++) {
            for (int w = 0; w < 5; w++) {
                cudaEventRecord(a);
                if (w == 0) cudaMemsetAsync(buf, 1, n);
                else if (w == 1) fill16<<<prop.multiProcessorCount * 8, 256>>>((uint4 *)buf, n / 16);
                else if (w == 2) fill16<<<prop.multiProcessorCount * 16, 512>>>((uint4 *)buf, n / 16);
                else if (w == 3) fill_bulk<16384><<<prop.multiProcessorCount * 8, 128, 16384>>>(buf, n / 16384);
                else fill_bulk<65536><<<prop.multiProcessorCount * 3, 128, 65536>>>(buf, n / 65536);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                float gbs = (float)(n / 1e9 / (ms / 1e3));
                if (rep >= 2 && gbs > best[w]) best[w] = gbs;
            }
        }
        cudaError_t e = cudaGetLastError();
        printf("  {\"bytes\": %zu, \"cudaMemsetAsync\": %.0f, \"stg128_grid8x256\": %.0f, \"stg128_grid16x512\": %.0f, \"tma_bulk_16k\": %.0f, \"tma_bulk_64k\": %.0f, \"err\": \"%s\"}%s\n",
               n, best[0], best[1], best[2], best[3], best[4], cudaGetErrorString(e), si < 3 ? "," : "");
    }
    printf("],\n \"per_game_pattern\": [\n");
    { // 128 KB per item (a 64x64 game's observations + masks), emitted as tiles of 1 .. 32 KB, one warp lane per item, waiting for completion per item
        size_t n = sizes[3];
        cudaFuncSetAttribute(fill_bulk_items<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024);
        cudaFuncSetAttribute(fill_bulk_items<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096);
        cudaFuncSetAttribute(fill_bulk_items<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
        cudaFuncSetAttribute(fill_bulk_items<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
        for (int wpsm = 8; wpsm <= 32; wpsm *= 2) {
            float best[4] = {0, 0, 0, 0};
            for (int rep = 0; rep < 8; rep++)
                for (int v = 0; v < 4; v++) {
                    int tile = v == 0 ? 1024 : v == 1 ? 4096 : v == 2 ? 16384 : 32768, per = 131072 / tile;
                    size_t items = n / 131072;
                    int ctas = prop.multiProcessorCount * (wpsm / 4);
                    cudaEventRecord(a);
                    if (v == 0) fill_bulk_items<1024><<<ctas, 128, 1024>>>(buf, items, per);
                    else if (v == 1) fill_bulk_items<4096><<<ctas, 128, 4096>>>(buf, items, per);
                    else if (v == 2) fill_bulk_items<16384><<<ctas, 128, 16384>>>(buf, items, per);
                    else fill_bulk_items<32768><<<ctas, 128, 32768>>>(buf, items, per);
                    cudaEventRecord(b); cudaEventSynchronize(b);
                    float ms; cudaEventElapsedTime(&ms, a, b);
                    float gbs = (float)(n / 1e9 / (ms / 1e3));
                    if (rep >= 2 && gbs > best[v]) best[v] = gbs;
                }
            printf("  {\"warps_per_sm\": %d, \"tile_1k\": %.0f, \"tile_4k\": %.0f, \"tile_16k\": %.0f, \"tile_32k\": %.0f, \"err\": \"%s\"}%s\n", wpsm, best[0], best[1], best[2], best[3],
                   cudaGetErrorString(cudaGetLastError()), wpsm < 32 ? "," : "");
        }
    }
    printf("]}\n");
    return 0;
}
