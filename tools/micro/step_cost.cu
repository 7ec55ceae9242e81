// What does one step of a register-tile right-looking sweep cost?  Variants of the loop body, fp32 / fp64, 128 / 512 threads.
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int VAR>
__global__ void __launch_bounds__(512) k(int n, long long* out, T* sink) {
    __shared__ T line[2][160];
    __shared__ T invd[80];
    const int tid = threadIdx.x;
    const int ti = tid / 17, tk = tid % 17, i0 = 4 * ti, c0 = 4 * tk;
    T m[4][4];
    for (int x = 0; x < 4; ++x) for (int y = 0; y < 4; ++y) m[x][y] = (T)(tid + x * 4 + y) * (T)1e-3;
    if (tid < 160) { line[0][tid] = (T)1 + (T)tid * (T)1e-3; line[1][tid] = (T)1; }
    if (tid < 80) invd[tid] = (T)1;
    __syncthreads();
    const long long t0 = clock64();
    for (int j = 0; j + 1 < n; ++j) {
        const T* ln = line[j & 1];
        T* nx = line[(j & 1) ^ 1];
        if (VAR >= 1) {
            if (i0 + 3 > j && c0 + 3 > j) {
                T ip;
                if (VAR == 1 || VAR == 4 || VAR == 5) ip = (T)1 / ln[j];
                else if (VAR == 2) ip = ln[j];
                else ip = invd[j];
                T f[4], cv[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) { f[x] = (i0 + x > j) ? ln[i0 + x] * ip : (T)0; cv[x] = (c0 + x > j) ? ln[c0 + x] : (T)0; }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) m[x][y] -= f[x] * cv[y];
                const int jn = j + 1;
                if (VAR != 5 && tk == (jn >> 2)) {
                    if (VAR == 4) {
#pragma unroll
                        for (int y = 0; y < 4; ++y) if ((jn & 3) == y) {
#pragma unroll
                            for (int x = 0; x < 4; ++x) if (i0 + x >= jn) nx[i0 + x] = m[x][y];
                        }
                    } else {
                        T v[4];
                        switch (jn & 3) {
                            case 0: v[0] = m[0][0]; v[1] = m[1][0]; v[2] = m[2][0]; v[3] = m[3][0]; break;
                            case 1: v[0] = m[0][1]; v[1] = m[1][1]; v[2] = m[2][1]; v[3] = m[3][1]; break;
                            case 2: v[0] = m[0][2]; v[1] = m[1][2]; v[2] = m[2][2]; v[3] = m[3][2]; break;
                            default: v[0] = m[0][3]; v[1] = m[1][3]; v[2] = m[2][3]; v[3] = m[3][3]; break;
                        }
#pragma unroll
                        for (int x = 0; x < 4; ++x) if (i0 + x >= jn) nx[i0 + x] = v[x];
                    }
                }
            }
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    T s = 0;
    for (int x = 0; x < 4; ++x) for (int y = 0; y < 4; ++y) s += m[x][y];
    sink[blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <typename T, int VAR>
void run(const char* nm, int n, int threads) {
    long long* d; T* sink; cudaMalloc(&d, 8); cudaMalloc(&sink, sizeof(T) * 148 * 512);
    for (int r = 0; r < 2; ++r) k<T, VAR><<<148, threads>>>(n, d, sink);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("%s n=%d threads=%d variant %d: %lld cycles = %.0f / step\n", nm, n, threads, VAR, c, (double)c / (n - 1));
}

int main() {
    printf("variants: 0 barrier only | 1 full (1/x) | 2 no reciprocal | 3 reciprocal precomputed in smem | 4 predicated publish (no switch) | 5 no publish\n");
    run<float, 0>("fp32", 33, 128); run<float, 1>("fp32", 33, 128); run<float, 2>("fp32", 33, 128); run<float, 3>("fp32", 33, 128); run<float, 4>("fp32", 33, 128); run<float, 5>("fp32", 33, 128);
    run<double, 0>("fp64", 33, 128); run<double, 1>("fp64", 33, 128); run<double, 2>("fp64", 33, 128); run<double, 3>("fp64", 33, 128); run<double, 4>("fp64", 33, 128); run<double, 5>("fp64", 33, 128);
    run<float, 0>("fp32", 65, 512); run<float, 1>("fp32", 65, 512); run<float, 2>("fp32", 65, 512); run<float, 4>("fp32", 65, 512); run<float, 5>("fp32", 65, 512);
    run<double, 0>("fp64", 65, 512); run<double, 1>("fp64", 65, 512); run<double, 2>("fp64", 65, 512); run<double, 4>("fp64", 65, 512); run<double, 5>("fp64", 65, 512);
    return 0;
}
