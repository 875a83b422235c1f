// k_util.cu -- data-parallel helpers around the wavefront machine: the normalisation constant b and the
// luminance CDF of the bootstrap (warp-shuffle reduce and scan), seed resampling, develop.
// Behavioural parity targets:
//   PathSampler::generateSeeds               src/libbidir/pathsampler.cpp:859-960
//   DiscreteDistribution                     include/mitsuba/core/pmf.h:60-200
//   DRMLTProcess::develop                    src/integrators/drmlt/drmlt_proc.cpp:813-854
#include "util_kernels.h"
#include "common.cuh"
#include "../../include/drmlt_b200.h"

// ---------------------------------------------------------------- b and the luminance CDF
// sum / count of the non-NaN luminances, bit-reproducible: fixed grid, fixed-order tree inside a block
// (warp shuffle, then warp 0 over the warp totals), then ONE thread adds the block partials in block order.
#define RED_BLOCKS (148 * 4)
__global__ void __launch_bounds__(256)
k_lum_reduce(const float *lum, long long n, double *partial /* [RED_BLOCKS][2] */) {
    __shared__ double ws[8], wc[8];
    double s = 0.0, c = 0.0;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const float v = lum[i];
        if (!isnan(v)) { s += (double) v; c += 1.0; }
    }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); c += __shfl_down_sync(0xffffffffu, c, o); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { ws[warp] = s; wc[warp] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tc = 0.0;
        for (int w = 0; w < 8; ++w) { ts += ws[w]; tc += wc[w]; }
        partial[2 * blockIdx.x] = ts; partial[2 * blockIdx.x + 1] = tc;
    }
}
__global__ void k_lum_final(const double *partial, double *out /* [2]: sum, count */) {
    double ts = 0.0, tc = 0.0;
    for (int b = 0; b < RED_BLOCKS; ++b) { ts += partial[2 * b]; tc += partial[2 * b + 1]; }
    out[0] = ts; out[1] = tc;
}

// Inclusive scan of max(lum,0) in double, three passes (block scan, scan of block totals, add).
#define SCAN_BLOCK 1024
#define SCAN_ITEMS 4
__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan_blocks(const float *lum, long long n, double *cdf /* n+1, cdf[0] = 0 */, double *blockSums,
              const float *depthWeight, unsigned long long first, int maxDepth) {
    __shared__ double warpSums[32];
    const long long base = (long long) blockIdx.x * SCAN_BLOCK * SCAN_ITEMS + (long long) threadIdx.x * SCAN_ITEMS;
    double v[SCAN_ITEMS], run = 0.0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const long long i = base + k;
        float x = i < n ? lum[i] : 0.f;
        if (!(x > 0.f) || isinf(x)) x = 0.f;     // NaN / negative / inf samples carry no seed
        double xd = (double) x;
        if (depthWeight && i < n) xd *= (double) depthWeight[(first + (unsigned long long) i) % (unsigned long long) maxDepth];
        run += xd; v[k] = run;
    }
    double incl = run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warpSums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        double w = warpSums[lane];
        for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        warpSums[lane] = w;
    }
    __syncthreads();
    const double offset = (incl - run) + (warp ? warpSums[warp - 1] : 0.0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { const long long i = base + k; if (i < n) cdf[i + 1] = v[k] + offset; }
    if (threadIdx.x == SCAN_BLOCK - 1) blockSums[blockIdx.x] = offset + run;
    if (blockIdx.x == 0 && threadIdx.x == 0) cdf[0] = 0.0;
}
__global__ void k_scan_sums(double *blockSums, int nb) {   // single block, serial over <= a few thousand entries per lane chunk
    __shared__ double carry;
    if (threadIdx.x == 0) carry = 0.0;
    __syncthreads();
    for (int start = 0; start < nb; start += blockDim.x) {
        const int i = start + threadIdx.x;
        double v = i < nb ? blockSums[i] : 0.0;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        __shared__ double ws[32];
        double incl = v;
        for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) ws[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            double w = lane < (blockDim.x >> 5) ? ws[lane] : 0.0;
            for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            ws[lane] = w;
        }
        __syncthreads();
        const double excl = incl - v + (warp ? ws[warp - 1] : 0.0) + carry;
        if (i < nb) blockSums[i] = excl;           // exclusive prefix of block totals
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
}
__global__ void k_scan_add(double *cdf, long long n, const double *blockSums) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i < n) cdf[i + 1] += blockSums[i / (SCAN_BLOCK * SCAN_ITEMS)];
}

// seedPDF.sample(next1D()) (pathsampler.cpp:946-954) on the un-normalised CDF
__global__ void k_resample(const double *cdf, long long n, unsigned long long seed, unsigned long long firstChain, int nChains,
                           unsigned long long bootFirst, int maxDepth, int technique,
                           unsigned long long *seedIdx, unsigned long long *chainId, int *depth) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nChains) return;
    const unsigned long long id = firstChain + (unsigned long long) c;
    const double total = cdf[n];
#ifdef DR_INDEPENDENT_RESAMPLING
    const double v = (double) keyed_uniform(seed, S_RESAMPLE, id, 0u, 0u) * total;
#else
    // stratified over the chains of this rank: chain c draws from the c-th of nChains equal slices of the CDF, so a seed of
    // relative luminance w starts floor(w nChains) or ceil(w nChains) chains instead of a binomially distributed number
    const double v = ((double) c + (double) keyed_uniform(seed, S_RESAMPLE, id, 0u, 0u)) / (double) nChains * total;
#endif
    long long lo = 0, hi = n + 1;
    while (lo < hi) { long long mid = (lo + hi) >> 1; if (cdf[mid] < v) lo = mid + 1; else hi = mid; }
    long long index = min(n - 1, max(0ll, lo - 1));
    while (index < n - 1 && cdf[index + 1] - cdf[index] == 0.0) ++index;
    const unsigned long long s = bootFirst + (unsigned long long) index;
    seedIdx[c] = s; chainId[c] = id;
    depth[c] = technique == DR_TECH_MMLT ? (int) (s % (unsigned long long) maxDepth) + 1 : -1;
}

// per-depth sums of the bootstrap luminances (MMLT: sample i has depth (first + i) % maxDepth + 1, pathsampler.cpp:886-890);
// out[d - 1] = sum.  maxDepth <= 32.  Fixed summation order (thread-strided, shared-memory tree, blocks in order), so a job
// with the same seed gets the same per-depth chain lengths on every run.
__global__ void __launch_bounds__(256)
k_depth_sums(const float *lum, long long n, unsigned long long first, int maxDepth, double *partial /* [RED_BLOCKS][32] */) {
    __shared__ double acc[256];
    // thread t of the grid only sees samples of ONE depth when its stride is a multiple of maxDepth: pad the stride
    const long long threads = (long long) gridDim.x * blockDim.x;
    const long long stride = (threads + maxDepth - 1) / maxDepth * maxDepth;
    const long long t = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    double s = 0.0;
    for (long long i = t; i < n; i += stride) {
        const float x = lum[i];
        if (x > 0.f && !isinf(x)) s += (double) x;
    }
    acc[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = 0.0;
        if ((int) threadIdx.x < maxDepth)
            for (int k = 0; k < 256; ++k)
                if ((int) ((first + (unsigned long long) (t - threadIdx.x + k)) % (unsigned long long) maxDepth) == (int) threadIdx.x) v += acc[k];
        partial[blockIdx.x * 32 + threadIdx.x] = v;
    }
}
__global__ void k_depth_final(const double *partial, int blocks, double *out) {
    double v = 0.0;
    for (int b = 0; b < blocks; ++b) v += partial[b * 32 + threadIdx.x];
    out[threadIdx.x] = v;
}
int depth_sums_scratch_doubles() { return RED_BLOCKS * 32; }
void launch_depth_sums(const float *lum, long long n, unsigned long long first, int maxDepth, double *scratch, double *out, cudaStream_t s) {
    k_depth_sums<<<RED_BLOCKS, 256, 0, s>>>(lum, n, first, maxDepth, scratch);
    k_depth_final<<<1, 32, 0, s>>>(scratch, RED_BLOCKS, out);
}

// ---------------------------------------------------------------- develop (drmlt_proc.cpp:813-854)
__global__ void k_film_luminance(const float4 *film, const float *importance, long long n, double *out) {
    double s = 0.0;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const float4 p = film[i];
        double l = (double) luminance(f3(p.x, p.y, p.z));
        if (importance) l *= (double) importance[i];         // avgLuminance += accum[i].getLuminance() * importanceMap[i] (:825-827)
        s += l;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}
__global__ void k_develop(const float4 *film, const float *importance, long long n, float factor, const float *direct, float *rgb) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = film[i];
    const float correction = importance ? factor * importance[i] : factor;                   // correction *= importanceMap[i] (:843-844)
    float r = p.x * correction, g = p.y * correction, b = p.z * correction;
    if (direct) { r += direct[3 * i]; g += direct[3 * i + 1]; b += direct[3 * i + 2]; }      // value += direct[i] (drmlt_proc.cpp:846-847)
    rgb[3 * i] = r; rgb[3 * i + 1] = g; rgb[3 * i + 2] = b;
}

// ---------------------------------------------------------------- two-stage MLT: importance map (src/libbidir/util.cpp:180-196)
// Developed first-stage image -> luminance (Spectrum::getLuminance, spectrum.h:640-650) -> separable up/down-sampling with
// precomputed, normalised filter taps (Resampler, include/mitsuba/core/rfilter.h:123-198; the tables are built on the
// host by dr_resample_luminance), clamped boundary (EClamp, :437-458), result clamped to [0, inf) after EACH pass
// (resampleAndClamp, :232-280).  One thread per output pixel.
__global__ void k_rgb_luminance(const float *rgb, long long n, double *lum) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i < n) lum[i] = (double) rgb[3 * i] * 0.212671f + (double) rgb[3 * i + 1] * 0.715160f + (double) rgb[3 * i + 2] * 0.072169f;
}
// pass along x: src [h][ws] -> dst [h][wt];  pass along y: src [hs][w] -> dst [ht][w]
__global__ void k_resample_axis(const double *src, int srcRes, int tgtRes, int other, int alongX, const int *start, const double *weights, int taps, double *dst) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i >= (long long) tgtRes * other) return;
    const int t = alongX ? (int) (i % tgtRes) : (int) (i / other), o = alongX ? (int) (i / tgtRes) : (int) (i % other);
    double result = 0.0;
    for (int j = 0; j < taps; ++j) {
        const int pos = min(max(start[t] + j, 0), srcRes - 1);
        const double v = alongX ? src[(size_t) o * srcRes + pos] : src[(size_t) pos * other + o];
        result += v * weights[(size_t) t * taps + j];
    }
    dst[alongX ? (size_t) o * tgtRes + t : (size_t) t * other + o] = fmax(0.0, result);       // min(max, max(min, result)), max = inf
}
__global__ void k_double_to_float(const double *src, long long n, float *dst) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (float) src[i];
}

// ---------------------------------------------------------------- launchers
int lum_reduce_scratch_doubles() { return 2 * RED_BLOCKS; }
void launch_lum_reduce(const float *lum, long long n, double *scratch, double *out, cudaStream_t s) {
    k_lum_reduce<<<RED_BLOCKS, 256, 0, s>>>(lum, n, scratch);
    k_lum_final<<<1, 1, 0, s>>>(scratch, out);
}
void launch_scan(const float *lum, long long n, double *cdf, double *blockSums, cudaStream_t s,
                 const float *depthWeight, unsigned long long first, int maxDepth) {
    const long long nb = scan_blocks(n);
    k_scan_blocks<<<(unsigned) nb, SCAN_BLOCK, 0, s>>>(lum, n, cdf, blockSums, depthWeight, first, maxDepth);
    k_scan_sums<<<1, 1024, 0, s>>>(blockSums, (int) nb);
    k_scan_add<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(cdf, n, blockSums);
}
long long scan_blocks(long long n) { return (n + SCAN_BLOCK * SCAN_ITEMS - 1) / (SCAN_BLOCK * SCAN_ITEMS); }
void launch_resample(const double *cdf, long long n, unsigned long long seed, unsigned long long firstChain, int nChains, unsigned long long bootFirst,
                     int maxDepth, int technique, unsigned long long *seedIdx, unsigned long long *chainId, int *depth, cudaStream_t s) {
    k_resample<<<(nChains + 127) / 128, 128, 0, s>>>(cdf, n, seed, firstChain, nChains, bootFirst, maxDepth, technique, seedIdx, chainId, depth);
}
void launch_film_luminance(const float4 *film, const float *importance, long long n, double *out, cudaStream_t s) { k_film_luminance<<<148 * 4, 256, 0, s>>>(film, importance, n, out); }
void launch_develop(const float4 *film, const float *importance, long long n, float factor, const float *direct, float *rgb, cudaStream_t s) {
    k_develop<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(film, importance, n, factor, direct, rgb);
}
void launch_rgb_luminance(const float *rgb, long long n, double *lum, cudaStream_t s) { k_rgb_luminance<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(rgb, n, lum); }
void launch_resample_axis(const double *src, int srcRes, int tgtRes, int other, int alongX, const int *start, const double *weights, int taps, double *dst, cudaStream_t s) {
    const long long n = (long long) tgtRes * other;
    k_resample_axis<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(src, srcRes, tgtRes, other, alongX, start, weights, taps, dst);
}
void launch_double_to_float(const double *src, long long n, float *dst, cudaStream_t s) { k_double_to_float<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(src, n, dst); }
