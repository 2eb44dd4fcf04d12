// heads_rate.cu -- what the value/action heads of the pair kernel cost per 256-env item when their weights come from shared
// memory (broadcast LDS.128, as shipped) or from the constant bank (a __grid_constant__ kernel parameter: FFMA with a c[][]
// operand, no shared-memory wavefronts).  Eight warps per CTA, each thread reduces 128 "columns" against 5 outputs, like
// warps 2..9 of k_forward_tc2; the activations are register values (the TMEM loads are left out on both sides).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/heads_rate.cu -o scripts/_build/heads_rate
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int H = 256;
struct HeadW { float w[5][H]; float b1[H]; };     // 6 KB

__global__ void __launch_bounds__(256, 1) k_heads_smem(const float* __restrict__ gw, float* out, long long* cyc, int items) {
    __shared__ __align__(16) float headw[5 * H];
    __shared__ __align__(16) float b1s[H];
    for (int i = threadIdx.x; i < 5 * H; i += blockDim.x) headw[i] = gw[i];
    for (int i = threadIdx.x; i < H; i += blockDim.x) b1s[i] = gw[5 * H + i];
    __syncthreads();
    const int chalf = (threadIdx.x >> 5) >> 2;
    float x = (float)threadIdx.x * 1e-3f, total = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < items; ++it) {
        float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
            const int col0 = chalf * 128 + q * 64;
#pragma unroll
            for (int j4 = 0; j4 < 16; ++j4) {
                const int col = col0 + 4 * j4;
                const float4 bb = *reinterpret_cast<const float4*>(b1s + col);
                float h[4] = {fmaxf(x + bb.x, 0.f), fmaxf(x * 1.1f + bb.y, 0.f), fmaxf(x * 1.2f + bb.z, 0.f), fmaxf(x * 1.3f + bb.w, 0.f)};
#pragma unroll
                for (int o = 0; o < 5; ++o) {
                    const float4 w = *reinterpret_cast<const float4*>(headw + o * H + col);
                    acc[o] = fmaf(h[0], w.x, acc[o]); acc[o] = fmaf(h[1], w.y, acc[o]);
                    acc[o] = fmaf(h[2], w.z, acc[o]); acc[o] = fmaf(h[3], w.w, acc[o]);
                }
            }
        }
        total += acc[0] + acc[1] + acc[2] + acc[3] + acc[4];
        x += total * 1e-9f;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = total;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void __launch_bounds__(256, 1) k_heads_const(const __grid_constant__ HeadW hw, float* out, long long* cyc, int items) {
    const int chalf = (threadIdx.x >> 5) >> 2;
    float x = (float)threadIdx.x * 1e-3f, total = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < items; ++it) {
        float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (chalf == 0) {
#pragma unroll
            for (int col = 0; col < 128; col += 4) {
                float h[4] = {fmaxf(x + hw.b1[col], 0.f), fmaxf(x * 1.1f + hw.b1[col + 1], 0.f), fmaxf(x * 1.2f + hw.b1[col + 2], 0.f), fmaxf(x * 1.3f + hw.b1[col + 3], 0.f)};
#pragma unroll
                for (int o = 0; o < 5; ++o) {
                    acc[o] = fmaf(h[0], hw.w[o][col], acc[o]); acc[o] = fmaf(h[1], hw.w[o][col + 1], acc[o]);
                    acc[o] = fmaf(h[2], hw.w[o][col + 2], acc[o]); acc[o] = fmaf(h[3], hw.w[o][col + 3], acc[o]);
                }
            }
        } else {
#pragma unroll
            for (int col = 128; col < 256; col += 4) {
                float h[4] = {fmaxf(x + hw.b1[col], 0.f), fmaxf(x * 1.1f + hw.b1[col + 1], 0.f), fmaxf(x * 1.2f + hw.b1[col + 2], 0.f), fmaxf(x * 1.3f + hw.b1[col + 3], 0.f)};
#pragma unroll
                for (int o = 0; o < 5; ++o) {
                    acc[o] = fmaf(h[0], hw.w[o][col], acc[o]); acc[o] = fmaf(h[1], hw.w[o][col + 1], acc[o]);
                    acc[o] = fmaf(h[2], hw.w[o][col + 2], acc[o]); acc[o] = fmaf(h[3], hw.w[o][col + 3], acc[o]);
                }
            }
        }
        total += acc[0] + acc[1] + acc[2] + acc[3] + acc[4];
        x += total * 1e-9f;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = total;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    const int items = 2000, grid = 148;
    HeadW* hw = new HeadW;
    for (int o = 0; o < 5; ++o) for (int i = 0; i < H; ++i) hw->w[o][i] = 0.01f * (float)((i * 7 + o) % 13 - 6);
    for (int i = 0; i < H; ++i) hw->b1[i] = 0.001f * (float)(i % 5);
    float *gw, *out; long long* cyc;
    cudaMalloc(&gw, sizeof(HeadW)); cudaMalloc(&out, sizeof(float) * grid * 256); cudaMalloc(&cyc, 8);
    cudaMemcpy(gw, hw, sizeof(HeadW), cudaMemcpyHostToDevice);
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
        k_heads_smem<<<grid, 256>>>(gw, out, cyc, items);
        cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("shared-memory weights : %8.1f cycles per item (8 warps, 128 columns x 5 outputs per thread)\n", (double)h / items);
        k_heads_const<<<grid, 256>>>(*hw, out, cyc, items);
        cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("constant-bank weights : %8.1f cycles per item   (%s)\n", (double)h / items, cudaGetErrorString(e));
    }
    return 0;
}
