// Micro-benchmark: issue rates of the integer / DPX / shuffle / LDS instructions the similarity kernels are built from.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o pipe_rates pipe_rates.cu
// Output: one JSON line per op: warp-instructions per clock per SM (so x32 = lane-ops/clk/SM).
#include <cstdio>
#include <cstdint>
#include <string>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

enum Op {
    OP_VIMAX3_U16X2, OP_VIMAX3_S16X2, OP_VIADDMAX_U16X2, OP_VIADDMAX_S16X2, OP_VIADDMAX_S16X2_RELU, OP_VIMAX_U16X2,
    OP_VIMAX3_S32, OP_VIADDMAX_S32, OP_IADD, OP_IADD3, OP_LOP3, OP_IMAD, OP_PRMT, OP_SHFL, OP_LDS32, OP_LDS128,
    OP_MIX_VIMAX3_IMAD, OP_MIX_VIMAX3_IADD, OP_MIX_VIMAX3_VIADDMAX, OP_MIX_LDS128_4VIMAX3_4IMAD, OP_VIMIN_U16X2_CMP,
    OP_IMAD_HI, OP_IMAD_WIDE, OP_SHF_FUNNEL, OP_IADD3_CARRY, OP_POPC, OP_HMNMX2, OP_MIX_VIADDMAX_HMNMX2, OP_COUNT
};
static const char *op_names[] = {
    "vimax3_u16x2", "vimax3_s16x2", "viaddmax_u16x2", "viaddmax_s16x2", "viaddmax_s16x2_relu", "vimax_u16x2",
    "vimax3_s32", "viaddmax_s32", "iadd", "iadd3", "lop3", "imad", "prmt", "shfl_up", "lds32", "lds128",
    "mix_vimax3+imad", "mix_vimax3+iadd", "mix_vimax3+viaddmax", "mix_lds128+4vimax3+4imad", "vimin_u16x2",
    "imad_hi", "imad_wide", "shf_funnel", "iadd3_carry_chain", "popc", "hmnmx2", "mix_viaddmax+hmnmx2", ""};
// instructions counted per inner iteration per chain
static const int op_instrs[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 9, 1, 1, 1, 1, 2, 1, 1, 2};

template <int OP>
__global__ void __launch_bounds__(256) bench(uint32_t *out, uint32_t seed, unsigned long long *cycles) {
    extern __shared__ uint4 smem[];
    uint32_t x[ILP], y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = seed * (threadIdx.x + 1) + i, y[i] = seed + 3 * i + threadIdx.x;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) smem[i] = make_uint4(i, i + 1, i + 2, seed);
    __syncthreads();
    uint32_t k = seed | 0x00010001u, m = seed ^ 0x5a5a;
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == OP_VIMAX3_U16X2) x[i] = __vimax3_u16x2(x[i], y[i], k);
            if (OP == OP_VIMAX3_S16X2) x[i] = __vimax3_s16x2(x[i], y[i], k);
            if (OP == OP_VIADDMAX_U16X2) x[i] = __viaddmax_u16x2(x[i], k, y[i]);
            if (OP == OP_VIADDMAX_S16X2) x[i] = __viaddmax_s16x2(x[i], k, y[i]);
            if (OP == OP_VIADDMAX_S16X2_RELU) x[i] = __viaddmax_s16x2_relu(x[i], k, y[i]);
            if (OP == OP_VIMAX_U16X2) x[i] = __vmaxu2(x[i] , y[i]) ^ k;
            if (OP == OP_VIMIN_U16X2_CMP) x[i] = __vimin3_u16x2(x[i], y[i], k);
            if (OP == OP_VIMAX3_S32) x[i] = __vimax3_s32(x[i], y[i], k);
            if (OP == OP_VIADDMAX_S32) x[i] = __viaddmax_s32(x[i], k, y[i]);
            if (OP == OP_IADD) x[i] = x[i] + y[i];
            if (OP == OP_IADD3) x[i] = x[i] + y[i] + k;
            if (OP == OP_LOP3) x[i] = (x[i] & y[i]) ^ k;
            if (OP == OP_IMAD) x[i] = x[i] * k + y[i];
            if (OP == OP_PRMT) x[i] = __byte_perm(x[i], y[i], k);
            if (OP == OP_SHFL) x[i] = __shfl_up_sync(0xffffffffu, x[i], 1) ;
            if (OP == OP_LDS32) x[i] = ((uint32_t *)smem)[(x[i] & 1023)];
            if (OP == OP_LDS128) { uint4 v = smem[(x[i] & 1023)]; x[i] = v.x ^ v.w; y[i] += v.y + v.z; }
            if (OP == OP_IMAD_HI) x[i] = __umulhi(x[i], k) ^ 0;
            if (OP == OP_IMAD_WIDE) { unsigned long long wide = (unsigned long long)x[i] * k + y[i]; x[i] = (uint32_t)wide; y[i] = (uint32_t)(wide >> 32); }
            if (OP == OP_SHF_FUNNEL) x[i] = __funnelshift_l(y[i], x[i], 1);
            if (OP == OP_IADD3_CARRY) { asm("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(k), "r"(m)); }
            if (OP == OP_POPC) x[i] = __popc(x[i]) + y[i];
            if (OP == OP_HMNMX2) { __half2 a = *reinterpret_cast<__half2 *>(&x[i]), b = *reinterpret_cast<__half2 *>(&y[i]); __half2 r = __hmax2(a, b); x[i] = *reinterpret_cast<uint32_t *>(&r) + 1; }
            if (OP == OP_MIX_VIADDMAX_HMNMX2) { y[i] = __viaddmax_u16x2(y[i], m, k); __half2 a = *reinterpret_cast<__half2 *>(&x[i]), b = *reinterpret_cast<__half2 *>(&y[i]); __half2 r = __hmax2(a, b); x[i] = *reinterpret_cast<uint32_t *>(&r); }
            if (OP == OP_MIX_VIMAX3_IMAD) { y[i] = y[i] * 1 + m; asm volatile("" : "+r"(y[i])); x[i] = __vimax3_u16x2(x[i], y[i], k); }
            if (OP == OP_MIX_VIMAX3_IADD) { y[i] = y[i] + m; x[i] = __vimax3_u16x2(x[i], y[i], k); }
            if (OP == OP_MIX_VIMAX3_VIADDMAX) { y[i] = __viaddmax_u16x2(y[i], m, k); x[i] = __vimax3_u16x2(x[i], y[i], k); }
        }
        if (OP == OP_MIX_LDS128_4VIMAX3_4IMAD) {
#pragma unroll
            for (int i = 0; i < ILP; i += 4) {
                uint4 v = smem[(threadIdx.x + it * 32 + i * 8) & 1023];
                uint32_t a = y[i] * k + v.x, b = y[i + 1] * k + v.y, c = y[i + 2] * k + v.z, d = y[i + 3] * k + v.w;
                x[i] = __vimax3_u16x2(x[i], a, k);
                x[i + 1] = __vimax3_u16x2(x[i + 1], b, x[i]);
                x[i + 2] = __vimax3_u16x2(x[i + 2], c, x[i + 1]);
                x[i + 3] = __vimax3_u16x2(x[i + 3], d, x[i + 2]);
            }
        }
    }
    unsigned long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int sms, int warps_per_sm, uint32_t *out, unsigned long long *cycles_d) {
    int threads = 256, blocks_per_sm = warps_per_sm * 32 / threads;
    int blocks = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    bench<OP><<<blocks, threads, 16384>>>(out, 12345u, cycles_d);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<OP><<<blocks, threads, 16384>>>(out, 12345u, cycles_d);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long *cyc = new unsigned long long[blocks];
    cudaMemcpy(cyc, cycles_d, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < blocks; ++i) mean += (double)cyc[i];
    mean /= blocks;
    int per_chain = op_instrs[OP];
    double warp_instrs_per_sm = (OP == OP_MIX_LDS128_4VIMAX3_4IMAD)
                                    ? (double)ITERS * (ILP / 4) * per_chain * warps_per_sm
                                    : (double)ITERS * ILP * per_chain * warps_per_sm;
    printf("{\"op\": \"%s\", \"warps_per_sm\": %d, \"warp_instr_per_clk_per_sm\": %.3f, \"mean_cycles\": %.0f, \"ms\": %.4f, "
           "\"eff_clock_mhz\": %.0f, \"err\": \"%s\"}\n",
           op_names[OP], warps_per_sm, warp_instrs_per_sm / mean, mean, ms, mean / (ms * 1e3),
           cudaGetErrorString(cudaGetLastError()));
    delete[] cyc;
}

template <int OP>
void sweep(int sms, uint32_t *out, unsigned long long *cyc) {
    if constexpr (OP < OP_COUNT) {
        run<OP>(sms, 16, out, cyc);
        run<OP>(sms, 32, out, cyc);
        sweep<OP + 1>(sms, out, cyc);
    }
}

/** `--dpx-peak`: sustained VIMNMX3.U16x2 issue rate over ~50 launches, the denominator of the int16-DPX roofline. */
static int dpx_peak() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int const sms = prop.multiProcessorCount, threads = 256, blocks = sms * 4;
    uint32_t *out;
    unsigned long long *cyc;
    cudaMalloc(&out, sizeof(uint32_t) * blocks * threads);
    cudaMalloc(&cyc, sizeof(unsigned long long) * blocks);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int i = 0; i < 5; ++i) bench<OP_VIMAX3_U16X2><<<blocks, threads, 16384>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    int const launches = 60;
    cudaEventRecord(e0);
    for (int i = 0; i < launches; ++i) bench<OP_VIMAX3_U16X2><<<blocks, threads, 16384>>>(out, 12345u, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double const warp_instrs = (double)launches * ITERS * ILP * (threads / 32) * blocks;
    double const lane_ops_per_s = warp_instrs * 32 / (ms * 1e-3);
    printf("{\"op\": \"vimax3_u16x2\", \"sms\": %d, \"ms\": %.3f, \"dpx_lane_ops_per_s\": %.4e, "
           "\"dpx_warp_instr_per_clk_per_sm_at_max_clock\": %.3f, \"max_clock_mhz\": %d, \"err\": \"%s\"}\n",
           sms, ms, lane_ops_per_s, warp_instrs / (ms * 1e-3) / sms / (prop.clockRate * 1e3), prop.clockRate / 1000,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}

int main(int argc, char **argv) {
    if (argc > 1 && std::string(argv[1]) == "--dpx-peak") return dpx_peak();
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    uint32_t *out;
    unsigned long long *cyc;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 8 * 256);
    cudaMalloc(&cyc, sizeof(unsigned long long) * sms * 8);
    sweep<0>(sms, out, cyc);
    return 0;
}
