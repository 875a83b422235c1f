// util_kernels.h -- host-callable launchers of k_util.cu
#pragma once
#include <cuda_runtime.h>
long long scan_blocks(long long n);
int lum_reduce_scratch_doubles();
void launch_lum_reduce(const float *lum, long long n, double *scratch, double *out /* [2]: sum, count */, cudaStream_t s);   // 2 launches, bit-reproducible
// 3 launches; depthWeight (optional, [maxDepth]): sample i is weighted by depthWeight[(first + i) % maxDepth]
void launch_scan(const float *lum, long long n, double *cdf /* n + 1 */, double *blockSums, cudaStream_t s,
                 const float *depthWeight = nullptr, unsigned long long first = 0, int maxDepth = 1);
int depth_sums_scratch_doubles();
void launch_depth_sums(const float *lum, long long n, unsigned long long first, int maxDepth, double *scratch, double *out /* [32] */, cudaStream_t s);   // 2 launches
void launch_resample(const double *cdf, long long n, unsigned long long seed, unsigned long long firstChain, int nChains, unsigned long long bootFirst,
                     int maxDepth, int technique, unsigned long long *seedIdx, unsigned long long *chainId, int *depth, cudaStream_t s);
void launch_film_luminance(const float4 *film, const float *importance /* or null */, long long n, double *out, cudaStream_t s);
void launch_develop(const float4 *film, const float *importance /* or null */, long long n, float factor, const float *direct /* or null */, float *rgb, cudaStream_t s);
// two-stage MLT importance map (util.cpp:180-196): luminance of an RGB image, one separable resampling pass, narrowing copy
void launch_rgb_luminance(const float *rgb, long long n, double *lum, cudaStream_t s);
void launch_resample_axis(const double *src, int srcRes, int tgtRes, int other, int alongX, const int *start, const double *weights, int taps, double *dst, cudaStream_t s);
void launch_double_to_float(const double *src, long long n, float *dst, cudaStream_t s);
