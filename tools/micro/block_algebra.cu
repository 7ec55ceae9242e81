// Cycle counts of the whole-CTA dense-algebra primitives of csrc/gp_block.cuh on one CTA per SM (debug aid, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I clip_gp_b200/csrc -I include tools/micro/block_algebra.cu -o gpurun_out/block_algebra && gpurun_out/block_algebra
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cstdint>
__device__ long long g_ts[4][16];
__device__ int g_ts_mode;
// the diagonal owner of block 2 (tile (2,2) = thread 5) and a far trailing tile record their clocks
#define BLK4_TS(i) do { if (blockIdx.x == 0) { if (threadIdx.x == 5) g_ts[0][i] = clock64(); if (threadIdx.x == 40) g_ts[1][i] = clock64(); } } while (0)
#include "gp_block.cuh"

using namespace clipgp::gp;

template <typename T>
__global__ void __launch_bounds__(512) k_chol(const T* A0, const T* B0, int n, int ncol, long long* cyc, T* outA, T* outB, int mode) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int ld = n | 1, ldb = ncol | 1;
    T* A = reinterpret_cast<T*>(sm);
    T* B = A + n * ld;
    T* invd = B + n * ldb;
    T* line = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(invd + 72) + 31) & ~(uintptr_t)31);
    __shared__ int flag;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) A[(i / n) * ld + i % n] = A0[i];
    for (int i = threadIdx.x; i < n * ncol; i += blockDim.x) B[(i / ncol) * ldb + i % ncol] = B0[i];
    __syncthreads();
    const long long t0 = clock64();
    bool f = false;
    if (mode == 0) f = cta_cholesky_solve<T, true>(A, n, ld, invd, B, ldb, ncol, line, &flag);
    else if (mode == 1) f = cta_cholesky_solve<T, false>(A, n, ld, invd, nullptr, 0, 0, line, &flag);
    else if (mode == 2) { f = block_cholesky_solve<T>(A, n, ld, invd, B, ldb, ncol, &flag); }
    else if (mode == 3) {
        if (threadIdx.x < 32) f = warp_cholesky<T>(A, n, ld, invd);
        __syncthreads();
        trsm_lower_left<T>(A, ld, invd, B, ldb, n, ncol);
    }
    const long long t1 = clock64();
    if (mode == 0 || mode == 2 || mode == 3) cta_trsm_lowerT_left<T>(A, ld, invd, B, ldb, n, ncol, line);   // -> (L L^T)^-1 B0
    const long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = f; cyc[3] = t0; }
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < n * n; i += blockDim.x) outA[i] = (i % n <= i / n) ? A[(i / n) * ld + i % n] : (T)0;
        for (int i = threadIdx.x; i < n * ncol; i += blockDim.x) outB[i] = B[(i / ncol) * ldb + i % ncol];
    }
}

template <typename T>
void run(const char* name, int n, int ncol, int threads) {
    std::vector<T> A(n * n), B(n * ncol), M(n * n);
    srand(1);
    for (auto& v : M) v = (T)(rand() / (double)RAND_MAX - 0.5);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = (i == j) ? 0.5 : 0.0; for (int k = 0; k < n; ++k) s += (double)M[i * n + k] * M[j * n + k]; A[i * n + j] = (T)s; }
    for (auto& v : B) v = (T)(rand() / (double)RAND_MAX - 0.5);
    T *dA, *dB, *oA, *oB; long long* dc;
    cudaMalloc(&dA, sizeof(T) * n * n); cudaMalloc(&dB, sizeof(T) * n * ncol); cudaMalloc(&oA, sizeof(T) * n * n); cudaMalloc(&oB, sizeof(T) * n * ncol); cudaMalloc(&dc, 64);
    cudaMemcpy(dA, A.data(), sizeof(T) * n * n, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), sizeof(T) * n * ncol, cudaMemcpyHostToDevice);
    const size_t smem = sizeof(T) * (n * (n | 1) + n * (ncol | 1) + 72 + 900) + 64;
    cudaFuncSetAttribute(k_chol<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode = 0; mode < 4; ++mode) {
        long long c[4] = {0, 0, 0, 0};
        for (int rep = 0; rep < 2; ++rep) {
            k_chol<T><<<148, threads, smem>>>(dA, dB, n, ncol, dc, oA, oB, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s mode %d: %s\n", name, mode, cudaGetErrorString(e)); return; }
        }
        cudaMemcpy(c, dc, 32, cudaMemcpyDeviceToHost);
        std::vector<T> X(n * ncol), Lh(n * n);
        cudaMemcpy(X.data(), oB, sizeof(T) * n * ncol, cudaMemcpyDeviceToHost); cudaMemcpy(Lh.data(), oA, sizeof(T) * n * n, cudaMemcpyDeviceToHost);
        double err = 0, errL = 0;
        for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int k = 0; k <= j; ++k) s += (double)Lh[i * n + k] * Lh[j * n + k]; errL = fmax(errL, fabs(s - A[i * n + j])); }
        if (mode != 1) for (int i = 0; i < n; ++i) for (int c2 = 0; c2 < ncol; ++c2) { double s = 0; for (int k = 0; k < n; ++k) s += (double)A[i * n + k] * X[k * ncol + c2]; err = fmax(err, fabs(s - B[i * ncol + c2])); }
        static const char* mn[] = {"blk4 chol+solve", "blk4 chol only", "block (smem) chol+solve", "one-warp chol + trsm"};
        if (mode < 2) {
            long long ts[4][16]; cudaMemcpyFromSymbol(ts, g_ts, sizeof(ts));
            for (int w = 0; w < 2; ++w) printf("      [thread %d] load+decode %lld | J=2: (a) %lld  bar %lld  (b) %lld  bar %lld  (c) %lld | whole loop %lld | write-back+pivots+scaling %lld\n", w ? 40 : 5,
                ts[w][0] - c[3], ts[w][2] - ts[w][1], ts[w][3] - ts[w][2], ts[w][4] - ts[w][3], ts[w][5] - ts[w][4], ts[w][6] - ts[w][5], ts[w][7] - ts[w][0], ts[w][8] - ts[w][7]);
        }
        printf("%-8s n=%d ncol=%d threads=%d  %-26s %8lld cycles | trsmT %8lld | fail=%lld  |LL^T-A|=%.2e |A X-B|=%.2e\n", name, n, ncol, threads, mn[mode], c[0], c[1], c[2], errL, err);
    }
}

int main() {
    run<double>("fp64", 65, 64, 512);
    run<float>("fp32", 64, 64, 512);
    run<double>("fp64", 33, 32, 128);
    run<float>("fp32", 32, 32, 128);
    return 0;
}
