// fp64_probe.cu — throughput of the instructions the true-cosine kernels lean on (B200): F2F.F64.F32, DFMA, and the
// integer re-biasing that replaces the conversion. Prints warp-instructions per clock per SM at 4 / 16 warps per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double cvt_int(float x) {
    const unsigned b = __float_as_uint(x);
    const unsigned hi = (b & 0x80000000u) | (((b & 0x7FFFFFFFu) >> 3) + 0x38000000u);
    return __hiloint2double((int)hi, (int)(b << 29));
}

template <int MODE>
__global__ void probe(const float* in, double* out, int iters, long long* cycles) {
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = in[threadIdx.x + 32 * i];
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {            // F2F + DADD-free: xor the bits so the conversion cannot be hoisted
                x[i] = __uint_as_float(__float_as_uint(x[i]) ^ (unsigned)it);
                acc[i] = __longlong_as_double(__double_as_longlong(acc[i]) ^ __double_as_longlong((double)x[i]));
            } else if (MODE == 1) {     // 8 independent DFMA chains
                acc[i] = fma(acc[i], 1.0000001, 0.5);
            } else if (MODE == 2) {     // integer conversion
                x[i] = __uint_as_float(__float_as_uint(x[i]) ^ (unsigned)it);
                acc[i] = __longlong_as_double(__double_as_longlong(acc[i]) ^ __double_as_longlong(cvt_int(x[i])));
            } else {                    // F2F feeding DFMA (the kernels' inner loop shape)
                x[i] = __uint_as_float(__float_as_uint(x[i]) ^ (unsigned)it);
                const double d = (double)x[i];
                acc[i] = fma(d, d, acc[i]);
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char* name, const float* in, double* out, long long* cyc, int sms) {
    for (int warps : {4, 16, 32}) {
        const int iters = 4096;
        probe<MODE><<<sms, warps * 32>>>(in, out, iters, cyc);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        const double winst = (double)iters * 8 * warps;
        printf("%-28s warps/SM %2d: %8.3f clk per warp-instr per SM  (%.2f lanes/clk/SM)\n", name, warps, c / winst, 32.0 * winst / c);
    }
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* in; double* out; long long* cyc;
    cudaMalloc(&in, 4096 * sizeof(float)); cudaMemset(in, 0x3c, 4096 * sizeof(float));
    cudaMalloc(&out, (size_t)sms * 1024 * sizeof(double)); cudaMalloc(&cyc, 8);
    run<0>("F2F.F64.F32 (+2 LOP)", in, out, cyc, sms);
    run<1>("DFMA x8 chains", in, out, cyc, sms);
    run<2>("integer f32->f64 (+2 LOP)", in, out, cyc, sms);
    run<3>("F2F + DFMA", in, out, cyc, sms);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
