// Throughput of u32 -> f32 conversion variants and of the six MUFU functions used per event.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t iters, float *out)
{
    uint32_t x = blockIdx.x * 256 + threadIdx.x, y = x * 2654435761u, z = x ^ 0x9E3779B9u;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (uint32_t i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u; y = y * 22695477u + 1u; z = z * 1103515245u + 12345u;   // 3 IMAD (FMA port)
        if (MODE == 0) {          // I2FP x3
            a0 += __uint2float_rn(x); a1 += __uint2float_rn(y); a2 += __uint2float_rn(z);
        } else if (MODE == 1) {   // bit trick x3: (w >> 9) | 0x3f800000
            a0 += __uint_as_float((x >> 9) | 0x3f800000u); a1 += __uint_as_float((y >> 9) | 0x3f800000u); a2 += __uint_as_float((z >> 9) | 0x3f800000u);
        } else if (MODE == 2) {   // nothing (baseline: 3 IMAD + 3 FADD of raw bits)
            a0 += __uint_as_float(x & 0x3fffffffu); a1 += __uint_as_float(y & 0x3fffffffu); a2 += __uint_as_float(z & 0x3fffffffu);
        } else if (MODE == 3) {   // 6 MUFU
            float f = __uint_as_float((x >> 9) | 0x3f800000u);
            a0 += __sinf(f) + __cosf(f); a1 += __log2f(f) + rsqrtf(f); float r, s; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f)); asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(f)); a2 += r + s;
        } else if (MODE == 4) {   // 6 MUFU + 3 I2FP
            float f = __uint2float_rn(x) * 2.3283064365386963e-10f + 1.0f, g = __uint2float_rn(y), h = __uint2float_rn(z);
            a0 += __sinf(f) + __cosf(f) + g; a1 += __log2f(f) + rsqrtf(f) + h; float r, s; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f)); asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(f)); a2 += r + s;
        }
    }
    out[blockIdx.x * 256 + threadIdx.x] = a0 + a1 + a2;
}

template <int MODE>
void run(const char *name)
{
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(100, out);
    cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(20000, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double per_s = 148.0 * 8 * 256 * 20000 / (ms * 1e-3);
    printf("%-40s %.1f cycles per warp-iteration per SMSP\n", name, 1.965e9 / (per_s / 32 / 592));
    cudaFree(out);
}

int main()
{
    run<2>("baseline: 3 IMAD + 3 LOP + 3 FADD");
    run<0>("3 IMAD + 3 I2FP + 3 FADD");
    run<1>("3 IMAD + 3 (SHF+LOP3) + 3 FADD");
    run<3>("6 MUFU (+ glue)");
    run<4>("6 MUFU + 3 I2FP (+ glue)");
    return 0;
}
