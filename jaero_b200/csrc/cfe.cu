// K2 — batched coarse frequency estimator.
//
// Replaces CoarseFreqEstimate::ProcessBasebandData (JAERO/coarsefreqestimate.cpp:90-137) for all
// channels of a batch: out=FFT(ring) -> zero bins [startbin,stopbin] (or raised-cosine window for
// 8400) -> in=N*IFFT(out) -> in=in^2 -> out=FFT(in) -> fftshift -> y=0.9y+0.1*10log10(max(|out|,1))
// -> fold search around the expected symbol-rate lines -> freq_offset_est, plus the
// emptyingcountdown gate (:134-135) and bigchange() (:84-88).
//
// B200 mapping: N = n1*n2 (128x128 for 2^14, 128x64 for 2^13) four-step FFT in double precision.
// The three transforms are fused into four memory passes by pairing the steps that work on the
// same row / column of the n1 x n2 matrix:
//   P1  column FFT (over r) of the linearised ring + twiddle                       ring -> A
//   P2  row FFT (over c) -> mask/window -> row IFFT + conj twiddle                 A    -> B
//   P3  column IFFT (over k1) -> square -> column FFT (over r) + twiddle           B    -> A
//   P4  row FFT (over c) -> |.| -> 10log10 -> smoothing into y (fft-shifted)       A    -> y
//   P5  fold search + emit gate (one warp per channel)
// Each pass moves 16-row / 16-column tiles (256 B segments) through shared memory; the 16 independent
// n<=128-point FFTs of a tile run as Stockham radix-8/4 passes with the butterflies in registers. FFT rounding differs from the
// CPU oracle's radix-2 (different factorisation) at the 1e-13 level; the bin decision is integer.
#include "demod_device.cuh"
#include <cooperative_groups.h>
#include <cstring>
#include <cstdio>
#include <cstdlib>

namespace jb {

static const int TILE = 16;
static const int MAXN = 128;
static const int CFE_THREADS = 256;

// ---- TILE independent length-n FFTs (n = 32, 64 or 128) held in natural order in s[f][.], natural order out.
// Stockham autosort with radix-8 / radix-4 butterflies kept in registers: 128 = 8*4*4, 64 = 8*8, 32 = 8*4, i.e. two or three
// passes over shared memory instead of log2(n) radix-2 passes. Every pass is "all threads read their inputs, barrier,
// compute + write, barrier" so it runs in place. tw = W_N^k table (N = big transform), tw_stride = N/n.
__device__ __forceinline__ double2 c_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 c_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 c_mul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <bool INV> __device__ __forceinline__ double2 c_rot(double2 a)      // multiply by -i (forward) / +i (inverse)
{ return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x); }

template <bool INV> __device__ __forceinline__ void dft4(double2 &a0, double2 &a1, double2 &a2, double2 &a3)
{
    const double2 t0 = c_add(a0, a2), t1 = c_sub(a0, a2), t2 = c_add(a1, a3), t3 = c_rot<INV>(c_sub(a1, a3));
    a0 = c_add(t0, t2); a1 = c_add(t1, t3); a2 = c_sub(t0, t2); a3 = c_sub(t1, t3);
}
template <bool INV> __device__ __forceinline__ void dft8(double2 *v)
{
    // even / odd split: X[k] = E[k] + W8^k O[k], X[k+4] = E[k] - W8^k O[k]
    double2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4<INV>(e0, e1, e2, e3);
    dft4<INV>(o0, o1, o2, o3);
    const double h = 0.70710678118654752440;
    // W8^1 = (1 -/+ i)/sqrt2, W8^2 = -/+ i, W8^3 = (-1 -/+ i)/sqrt2   (upper sign: forward)
    const double2 w1 = INV ? make_double2(h, h) : make_double2(h, -h);
    const double2 w3 = INV ? make_double2(-h, h) : make_double2(-h, -h);
    o1 = c_mul(o1, w1); o2 = c_rot<INV>(o2); o3 = c_mul(o3, w3);
    v[0] = c_add(e0, o0); v[4] = c_sub(e0, o0);
    v[1] = c_add(e1, o1); v[5] = c_sub(e1, o1);
    v[2] = c_add(e2, o2); v[6] = c_sub(e2, o2);
    v[3] = c_add(e3, o3); v[7] = c_sub(e3, o3);
}

template <bool INV, int R>
__device__ __forceinline__ void stockham_pass(double2 (*s)[MAXN + 1], int n, int Ns, const double2 *__restrict__ tw, int tw_stride)
{
    constexpr int PER = 8 / R;                       // butterflies per thread per round (8 complex values in registers)
    const int nb = n / R;                            // butterflies per sequence
    const int total = TILE * nb;
    const int wmul = tw_stride * (n / (Ns * R));     // table step of exp(-2 pi i /(Ns R))
    for (int base = 0; base < total; base += CFE_THREADS * PER) {
        double2 v[8];
        int ff[PER], jj[PER];
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int b = base + threadIdx.x + q * CFE_THREADS;
            ff[q] = -1;
            if (b < total) {
                const int f = b / nb, j = b - f * nb;
                ff[q] = f; jj[q] = j;
#pragma unroll
                for (int t = 0; t < R; t++) v[q * R + t] = s[f][j + t * nb];
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER; q++) {
            if (ff[q] >= 0) {
                const int j = jj[q], k = j % Ns;
#pragma unroll
                for (int t = 1; t < R; t++) {
                    double2 w = tw[(t * k * wmul)];
                    if (INV) w.y = -w.y;
                    v[q * R + t] = c_mul(v[q * R + t], w);
                }
                if (R == 8) dft8<INV>(&v[q * R]);
                else dft4<INV>(v[q * R + 0], v[q * R + 1], v[q * R + 2], v[q * R + 3]);
                const int ob = (j / Ns) * Ns * R + k;
#pragma unroll
                for (int t = 0; t < R; t++) s[ff[q]][ob + t * Ns] = v[q * R + t];
            }
        }
        __syncthreads();
    }
}

template <bool INV>
__device__ __forceinline__ void tile_fft(double2 (*s)[MAXN + 1], int n, const double2 *__restrict__ tw, int tw_stride)
{
    stockham_pass<INV, 8>(s, n, 1, tw, tw_stride);
    if (n == 128) { stockham_pass<INV, 4>(s, n, 8, tw, tw_stride); stockham_pass<INV, 4>(s, n, 32, tw, tw_stride); }
    else if (n == 64) stockham_pass<INV, 8>(s, n, 8, tw, tw_stride);
    else stockham_pass<INV, 4>(s, n, 8, tw, tw_stride);          // n == 32
}

// P1 / P3: column pass. grid = (n2/TILE, channels)
template <bool FUSED_INV_SQUARE>
__global__ void __launch_bounds__(CFE_THREADS)
cfe_col_kernel(CfePlan pl, const double2 *__restrict__ src, double2 *__restrict__ dst, size_t src_pitch, int rot, int ring_len, int ch0)
{
    __shared__ double2 s[TILE][MAXN + 1];
    const int ch = blockIdx.y;
    const int c0 = blockIdx.x * TILE;
    const int n1 = pl.n1, n2 = pl.n2, N = pl.nfft;
    const double2 *in = src + (size_t)(ch0 + ch) * src_pitch;
    double2 *out = dst + (size_t)ch * N;
    // load column tile: element (r, c0+cc) of the n1 x n2 matrix, n = n2*r + c  (rot linearises the ring:
    // bbtmpbuff[j] = bbcycbuff[(ptr+j)%N], oqpskdemodulator.cpp:418-424)
    for (int e = threadIdx.x; e < n1 * TILE; e += CFE_THREADS) {
        const int r = e / TILE, cc = e - r * TILE;
        int n = n2 * r + c0 + cc;
        if (!FUSED_INV_SQUARE) { n += rot; if (n >= ring_len) n -= ring_len; }
        s[cc][r] = in[n];
    }
    __syncthreads();
    if (FUSED_INV_SQUARE) {
        // column IFFT over k1 -> x'[n2*r+c]; square; then forward again
        tile_fft<true>(s, n1, pl.tw, N / n1);
        // square in place (natural order in, natural order out of the Stockham passes)
        for (int e = threadIdx.x; e < n1 * TILE; e += CFE_THREADS) {
            const int cc = e / n1, r = e - cc * n1;
            const double2 x = s[cc][r];
            s[cc][r] = make_double2(x.x * x.x - x.y * x.y, x.x * x.y + x.y * x.x);   // in[i]*in[i] (:103)
        }
        __syncthreads();
    }
    tile_fft<false>(s, n1, pl.tw, N / n1);
    // twiddle W_N^{c*k1} and store A[k1][c]
    for (int e = threadIdx.x; e < n1 * TILE; e += CFE_THREADS) {
        const int k1 = e / TILE, cc = e - k1 * TILE;
        const int c = c0 + cc;
        const double2 w = pl.tw[(c * k1) & (N - 1)];
        const double2 x = s[cc][k1];
        out[(size_t)k1 * n2 + c] = make_double2(x.x * w.x - x.y * w.y, x.x * w.y + x.y * w.x);
    }
}

// P2: row pass, forward -> mask -> inverse -> conj twiddle. grid = (n1/TILE, channels)
__global__ void __launch_bounds__(CFE_THREADS)
cfe_row_mask_kernel(CfePlan pl, const double2 *__restrict__ src, double2 *__restrict__ dst)
{
    __shared__ double2 s[TILE][MAXN + 1];
    const int ch = blockIdx.y;
    const int r0 = blockIdx.x * TILE;
    const int n1 = pl.n1, n2 = pl.n2, N = pl.nfft;
    const double2 *in = src + (size_t)ch * N;
    double2 *out = dst + (size_t)ch * N;
    for (int e = threadIdx.x; e < n2 * TILE; e += CFE_THREADS) {
        const int rr = e / n2, c = e - rr * n2;
        s[rr][c] = in[(size_t)(r0 + rr) * n2 + c];
    }
    __syncthreads();
    tile_fft<false>(s, n2, pl.tw, N / n2);
    // X[k1 + n1*k2] sits at s[k1-r0][k2]; mask (:99-100) in place
    for (int e = threadIdx.x; e < n2 * TILE; e += CFE_THREADS) {
        const int rr = e / n2, k2 = e - rr * n2;
        const int k = (r0 + rr) + n1 * k2;
        double2 x = s[rr][k2];
        if (!pl.is8400) { if (k >= pl.startbin && k <= pl.stopbin) x = make_double2(0.0, 0.0); }
        else { const double w = pl.window[k]; x = make_double2(x.x * w, x.y * w); }
        s[rr][k2] = x;
    }
    __syncthreads();
    tile_fft<true>(s, n2, pl.tw, N / n2);
    for (int e = threadIdx.x; e < n2 * TILE; e += CFE_THREADS) {
        const int rr = e / n2, c = e - rr * n2;
        const int k1 = r0 + rr;
        double2 w = pl.tw[(c * k1) & (N - 1)];
        w.y = -w.y;
        const double2 x = s[rr][c];
        out[(size_t)k1 * n2 + c] = make_double2(x.x * w.x - x.y * w.y, x.x * w.y + x.y * w.x);
    }
}

// P4: row pass, forward -> |.| -> log -> smoothing. grid = (n1/TILE, channels)
__global__ void __launch_bounds__(CFE_THREADS)
cfe_row_logmag_kernel(CfePlan pl, DemodParams p, const double2 *__restrict__ src, int ch0)
{
    __shared__ double2 s[TILE][MAXN + 1];
    const int ch = blockIdx.y;
    const int r0 = blockIdx.x * TILE;
    const int n1 = pl.n1, n2 = pl.n2, N = pl.nfft;
    const double2 *in = src + (size_t)ch * N;
    double *y = pl.y + (size_t)(ch0 + ch) * N;
    const bool bigchange = p.I[(size_t)I_ZERO_BB * p.cpad + ch0 + ch] != 0;     // y[i]=20 pending (coarsefreqestimate.cpp:87)
    for (int e = threadIdx.x; e < n2 * TILE; e += CFE_THREADS) {
        const int rr = e / n2, c = e - rr * n2;
        s[rr][c] = in[(size_t)(r0 + rr) * n2 + c];
    }
    __syncthreads();
    tile_fft<false>(s, n2, pl.tw, N / n2);
    // Y[k1 + n1*k2]; fftshift (:105): shifted index i = (k + N/2) % N = k1 + n1*((k2 + n2/2) % n2)
    for (int e = threadIdx.x; e < n2 * TILE; e += CFE_THREADS) {
        const int k2 = e / TILE, rr = e - k2 * TILE;         // rr fastest -> 16 consecutive i per k2
        const int k1 = r0 + rr;
        const int k2s = (k2 + (n2 >> 1)) & (n2 - 1);
        const int i = k1 + n1 * k2s;
        const double2 x = s[rr][k2];
        const double mag = hypot(x.x, x.y);
        const double yo = bigchange ? 20.0 : y[i];
        y[i] = yo * 0.9 + 0.1 * 10 * log10(fmax(mag, 1.0));                    // :108
    }
}

// P5: fold search (:112-131) + emit gate (:134-135). One warp per channel.
__global__ void __launch_bounds__(128)
cfe_search_kernel(CfePlan pl, DemodParams p)
{
    const int lane = threadIdx.x & 31;
    const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ch >= p.n_channels) return;
    const double *y = pl.y + (size_t)ch * pl.nfft;
    const int N = pl.nfft, epb = pl.expectedpeakbin;
    double best = 0.0; int besti = 0x7fffffff;
    int i0 = pl.lo;
    if (pl.lo - epb - 1 >= 0 && pl.hi + epb + 1 < N) {
        // every index of the fold is inside the spectrum (true for all four rates): no per-bin range tests, and four
        // candidate bins per lane and round, their 24 loads requested together (the loop is load-latency bound otherwise)
        for (; i0 + 128 <= pl.hi; i0 += 128) {
            double a[4][3], c[4][3];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + lane + 32 * u;
#pragma unroll
                for (int j = -1; j <= 1; j++) { a[u][j + 1] = y[i - epb - j]; c[u][j + 1] = y[i + epb + j]; }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                double val = 0;
#pragma unroll
                for (int j = 0; j < 3; j++) val += (a[u][j] + c[u][j]);
                if (val > best) { best = val; besti = i0 + lane + 32 * u; }   // ascending i per lane: the first maximum wins
            }
        }
    }
    for (int i = i0 + lane; i < pl.hi; i += 32) {
        if ((i < 0) || (i >= N)) continue;
        double val = 0;
        for (int j = -1; j <= 1; j++) {
            if (((i - epb - j) < 0) || ((i + epb + j) >= N)) continue;
            val += (y[i - epb - j] + y[i + epb + j]);
        }
        if (val > best) { best = val; besti = i; }           // strict >: the first maximum wins
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, off);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if (lane == 0) {
        const int zmaxloc = (best > 0.0) ? besti : N / 2;    // zmax starts at 0, zmaxloc at nfft/2
        const double est = -((double)(zmaxloc - N / 2)) * pl.hzperbin * 0.5;   // :131
        p.D[(size_t)D_CFE_EST * p.cpad + ch] = est;
        int &emptying = p.I[(size_t)I_EMPTYING * p.cpad + ch];
        if (emptying <= 0) p.cfe_est_out[ch] = est;
        else { emptying--; p.cfe_est_out[ch] = 0.0; }
        p.I[(size_t)I_ZERO_BB * p.cpad + ch] = 0;
    }
}

// ================================================================================================ cluster-resident estimator
// nfft = 16384 (both OQPSK modes). One thread-block cluster of 8 CTAs owns one channel at a time and keeps the whole
// 128 x 128 working matrix (256 KB of complex doubles) in the distributed shared memory of its CTAs through all three
// transforms: CTA q holds 16 columns (column passes) or 16 rows (row passes); the three layout changes are pulls from
// the peers' shared memory (DSMEM) with the four-step twiddle folded into the pull. A CTA needs 74 KB of shared memory
// and 256 threads, so three clusters' CTAs share an SM and three channels are in flight per SM: one channel's barrier /
// pull latency is covered by the others' butterflies. HBM traffic per channel falls to the ring read (256 KB) plus the
// read-modify-write of y (2 x 128 KB), from eight 256 KB matrix passes before.
namespace cgx = cooperative_groups;

static const int CC_CL = 8, CC_T = 256, CC_SEQ = 16, CC_RS = 143;             // cluster size, threads, sequences per CTA, row stride
static const int CC_BUF = CC_SEQ * CC_RS * 16;                                // one working buffer (bytes)
static const int CC_TW2 = 128, CC_TW3 = 152;                                  // offsets (in double2) of the per-pass twiddle copies
static const int CC_SM_TOTAL = 2 * CC_BUF + (128 + 24 + 32) * 16;

__device__ __forceinline__ int cc_ph(int e) { return e + (e >> 3); }          // padded position inside a 128-point sequence

// log10 for finite x >= 1 (the argument is max(|X|^2, 1)): fdlibm's log kernel — x = 2^k m, m in [sqrt(1/2), sqrt(2)),
// s = (m-1)/(m+1), degree-14 polynomial in s — with the quotient by div_fast and the final scaling split as in fdlibm's
// e_log10.c. Within 2 ulp of glibc's log10 on 5 M arguments in [1, 1e40] (host twin: tools/micro/log10_test.c; the library call it
// replaces is itself only specified to 1 ulp), about half the instructions; the smoothed spectrum feeds an arg-max over
// sums of six bins, where differences of that size cannot matter unless the library's own rounding would.
__device__ __forceinline__ double log10_ge1(double x)
{
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                 Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
    const double ivln10 = 4.34294481903251816668e-01, log10_2hi = 3.01029995663611771306e-01, log10_2lo = 3.69423907715893078616e-13;
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int i = (hi + 0x95f64) & 0x100000;                 // m >= sqrt(2): halve it
    hi |= (i ^ 0x3ff00000);
    k += (i >> 20);
    const double f = __hiloint2double(hi, lo) - 1.0;
    const double s = div_fast(f, 2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    const double hfsq = 0.5 * f * f;
    const double lg = f - (hfsq - s * (hfsq + (t2 + t1)));
    const double dk = (double)k;
    return (dk * log10_2lo + ivln10 * lg) + dk * log10_2hi;
}

// 16 x FFT-128 in place (same factorisation and arithmetic as tile_fft for n = 128: Stockham radix 8, 4, 4). Warp w owns
// sequences w and w+8 through all three passes, so the passes are ordered by __syncwarp() only: each pass reads its
// butterflies' inputs into registers, __syncwarp, writes the outputs back into the same rows. `last` receives
// (sequence, position, value) of the final pass and normally stores it back (mask / square are fused there).
// FIRST = false: the caller has already run the first (radix-8) pass from registers — the values it pulled from the ring or
// from its peers are exactly one butterfly's inputs — stored the outputs (row[8j + t]) and passed a __syncthreads().
template <bool INV, bool FIRST, class Store>
__device__ __forceinline__ void cc_fft(double2 *__restrict__ buf, const double2 *__restrict__ tws, Store last)
{
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (FIRST) {   // radix 8, Ns = 1: lanes 0-15 -> sequence w, lanes 16-31 -> sequence w+8
        double2 *row = buf + (w + ((l & 16) >> 1)) * CC_RS;
        const int j = l & 15;
        double2 v[8];
#pragma unroll
        for (int t = 0; t < 8; t++) v[t] = row[cc_ph(j + 16 * t)];
        __syncwarp();
        dft8<INV>(v);
#pragma unroll
        for (int t = 0; t < 8; t++) row[cc_ph(8 * j + t)] = v[t];
        __syncwarp();
    }
    double2 *r0 = buf + w * CC_RS, *r1 = buf + (w + 8) * CC_RS;
    // The twiddles of a pass depend on the lane only. Read straight from the 128-entry table the second pass's strides
    // (4k, 8k, 12k) put a quarter-warp's eight 16-byte loads on two, one and two bank groups (4-, 8- and 4-way conflicts) and
    // the third pass's 2l on four (2-way); the kernel keeps compact copies instead: tws[CC_TW2 + 8(m-1) + k] = W^(4mk),
    // tws[CC_TW3 + l] = W^(2l). Same table values, so the arithmetic is unchanged.
    auto radix4 = [&](double2 &v0, double2 &v1, double2 &v2, double2 &v3, double2 w1, double2 w2, double2 w3) {
        if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
        v1 = c_mul(v1, w1); v2 = c_mul(v2, w2); v3 = c_mul(v3, w3);
        dft4<INV>(v0, v1, v2, v3);
    };
    {   // radix 4, Ns = 8 (both sequences, j = lane)
        double2 a0 = r0[cc_ph(l)], a1 = r0[cc_ph(l + 32)], a2 = r0[cc_ph(l + 64)], a3 = r0[cc_ph(l + 96)];
        double2 b0 = r1[cc_ph(l)], b1 = r1[cc_ph(l + 32)], b2 = r1[cc_ph(l + 64)], b3 = r1[cc_ph(l + 96)];
        __syncwarp();
        const int k = l & 7, ob = (l >> 3) * 32 + k;
        const double2 w1 = tws[CC_TW2 + k], w2 = tws[CC_TW2 + 8 + k], w3 = tws[CC_TW2 + 16 + k];
        radix4(a0, a1, a2, a3, w1, w2, w3); radix4(b0, b1, b2, b3, w1, w2, w3);
        r0[cc_ph(ob)] = a0; r0[cc_ph(ob + 8)] = a1; r0[cc_ph(ob + 16)] = a2; r0[cc_ph(ob + 24)] = a3;
        r1[cc_ph(ob)] = b0; r1[cc_ph(ob + 8)] = b1; r1[cc_ph(ob + 16)] = b2; r1[cc_ph(ob + 24)] = b3;
        __syncwarp();
    }
    {   // radix 4, Ns = 32
        double2 a0 = r0[cc_ph(l)], a1 = r0[cc_ph(l + 32)], a2 = r0[cc_ph(l + 64)], a3 = r0[cc_ph(l + 96)];
        double2 b0 = r1[cc_ph(l)], b1 = r1[cc_ph(l + 32)], b2 = r1[cc_ph(l + 64)], b3 = r1[cc_ph(l + 96)];
        __syncwarp();
        const double2 w1 = tws[l], w2 = tws[CC_TW3 + l], w3 = tws[3 * l];
        radix4(a0, a1, a2, a3, w1, w2, w3); radix4(b0, b1, b2, b3, w1, w2, w3);
        last(w, l, a0); last(w, l + 32, a1); last(w, l + 64, a2); last(w, l + 96, a3);
        last(w + 8, l, b0); last(w + 8, l + 32, b1); last(w + 8, l + 64, b2); last(w + 8, l + 96, b3);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(CC_T, 2)
cfe_cluster_kernel(CfePlan pl, DemodParams p, int oldest)
{
    extern __shared__ __align__(128) unsigned char cc_smem[];
    double2 *bufA = reinterpret_cast<double2 *>(cc_smem);
    double2 *bufB = reinterpret_cast<double2 *>(cc_smem + CC_BUF);
    double2 *tws = reinterpret_cast<double2 *>(cc_smem + 2 * CC_BUF);
    cgx::cluster_group cluster = cgx::this_cluster();
    const int q = (int)cluster.block_rank();
    const int n_clusters = gridDim.x / CC_CL, cid = blockIdx.x / CC_CL;
    const int N = 16384, ring_len = p.bb_len;
    const double2 *__restrict__ twN = pl.tw;
    if (threadIdx.x < 128) tws[threadIdx.x] = twN[threadIdx.x * (N / 128)];
    else if (threadIdx.x < 128 + 24) { const int e = threadIdx.x - 128, m = e >> 3, k = e & 7; tws[CC_TW2 + e] = twN[(4 * (m + 1) * k) * (N / 128)]; }
    else if (threadIdx.x < 128 + 24 + 32) { const int l = threadIdx.x - 152; tws[CC_TW3 + l] = twN[(2 * l) * (N / 128)]; }
    // The four-step twiddles W_N^(c*k1) are read from the same table the reference-order transforms use (a product of two
    // smaller tables breaks the exact conjugate symmetry of the table and with it the estimator's tie-breaks on symmetric
    // spectra). Their indices do not depend on the data, so the loads are issued ahead of the cluster barrier they follow.
    __syncthreads();
    // peers' buffers
    const double2 *rA[CC_CL], *rB[CC_CL];
#pragma unroll
    for (int s = 0; s < CC_CL; s++) { rA[s] = cluster.map_shared_rank(bufA, s); rB[s] = cluster.map_shared_rank(bufB, s); }
    bool arrived = false;
    cluster.sync();                        // every CTA of the cluster is resident before the first remote access
    for (int ch = cid; ch < p.n_channels; ch += n_clusters) {
        // ---- this CTA's 16 columns of the linearised ring (oqpskdemodulator.cpp:418-424) -> A[cc][r]: 256 B runs per r
        double2 colv[8];
        {
            const double2 *ring = p.bb + (size_t)ch * ring_len;
            const int cc = threadIdx.x & 15, r0i = threadIdx.x >> 4;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                int n = oldest + 128 * (r0i + 16 * i) + 16 * q + cc;
                if (n >= ring_len) n -= ring_len;
                if (n >= ring_len) n -= ring_len;
                colv[i] = ring[n];
            }
        }
        // ---- P1: column FFT over r                                                   A[cc][k1]
        // this thread's eight samples r = r0i + 16 i of column cc are butterfly r0i of the first pass: it runs from registers
        dft8<false>(colv);
        if (arrived) { cluster.barrier_wait(); arrived = false; }   // the peers have pulled the previous channel's P3 result out of A
        {
            const int cc = threadIdx.x & 15, r0i = threadIdx.x >> 4;
#pragma unroll
            for (int i = 0; i < 8; i++) bufA[cc * CC_RS + cc_ph(8 * r0i + i)] = colv[i];
        }
        __syncthreads();
        cc_fft<false, false>(bufA, tws, [&](int f, int e, double2 v) { bufA[f * CC_RS + cc_ph(e)] = v; });
        double2 tw8[8];
        {
            const int kk = threadIdx.x & 15, cc = threadIdx.x >> 4;
#pragma unroll
            for (int s = 0; s < 8; s++) tw8[s] = __ldg(&twN[((16 * s + cc) * (16 * q + kk)) & (N - 1)]);
        }
        cluster.sync();
        // rows k1 = 16q+kk, all c: B[kk][c] = A_src[cc][k1] * W_N^{c k1}   (kk fastest: contiguous remote reads)
        {
            const int kk = threadIdx.x & 15, cc = threadIdx.x >> 4;
            double2 x[8];
#pragma unroll
            for (int s = 0; s < 8; s++) x[s] = c_mul(rA[s][cc * CC_RS + cc_ph(16 * q + kk)], tw8[s]);
            dft8<false>(x);                // c = cc + 16 s: butterfly cc of row kk's first pass
#pragma unroll
            for (int s = 0; s < 8; s++) bufB[kk * CC_RS + cc_ph(8 * cc + s)] = x[s];
        }
        __syncthreads();
        // ---- P2: row FFT over c -> mask (:99-100) -> row IFFT, in place               B[kk][c]
        cc_fft<false, false>(bufB, tws, [&](int f, int e, double2 v) {
            const int k = (16 * q + f) + 128 * e;          // X[k1 + n1*k2]
            if (!pl.is8400) { if (k >= pl.startbin && k <= pl.stopbin) v = make_double2(0.0, 0.0); }
            else { const double w = pl.window[k]; v = make_double2(v.x * w, v.y * w); }
            bufB[f * CC_RS + cc_ph(e)] = v;
        });
        cc_fft<true, true>(bufB, tws, [&](int f, int e, double2 v) { bufB[f * CC_RS + cc_ph(e)] = v; });
        {
            const int cc = threadIdx.x & 15, kk = threadIdx.x >> 4;
#pragma unroll
            for (int s = 0; s < 8; s++) tw8[s] = __ldg(&twN[((16 * q + cc) * (16 * s + kk)) & (N - 1)]);
        }
        cluster.sync();                    // every peer has finished reading A (it passed the pull above before its own P2)
        // columns c = 16q+cc, all k1: A[cc][k1] = B_src[kk][c] * conj(W_N^{c k1})   (cc fastest: contiguous remote reads)
        {
            const int cc = threadIdx.x & 15, kk = threadIdx.x >> 4;
            double2 x[8];
#pragma unroll
            for (int s = 0; s < 8; s++) {
                double2 w = tw8[s]; w.y = -w.y;
                x[s] = c_mul(rB[s][kk * CC_RS + cc_ph(16 * q + cc)], w);
            }
            dft8<true>(x);                 // k1 = kk + 16 s: butterfly kk of column cc's first pass
#pragma unroll
            for (int s = 0; s < 8; s++) bufA[cc * CC_RS + cc_ph(8 * kk + s)] = x[s];
        }
        __syncthreads();
        // ---- P3: column IFFT over k1 -> square (:103) -> column FFT over r, in place  A[cc][k1]
        cc_fft<true, false>(bufA, tws, [&](int f, int e, double2 v) {
            bufA[f * CC_RS + cc_ph(e)] = make_double2(v.x * v.x - v.y * v.y, v.x * v.y + v.y * v.x);
        });
        cc_fft<false, true>(bufA, tws, [&](int f, int e, double2 v) { bufA[f * CC_RS + cc_ph(e)] = v; });
        {
            const int kk = threadIdx.x & 15, cc = threadIdx.x >> 4;
#pragma unroll
            for (int s = 0; s < 8; s++) tw8[s] = __ldg(&twN[((16 * s + cc) * (16 * q + kk)) & (N - 1)]);
        }
        cluster.sync();
        {
            const int kk = threadIdx.x & 15, cc = threadIdx.x >> 4;
            double2 x[8];
#pragma unroll
            for (int s = 0; s < 8; s++) x[s] = c_mul(rA[s][cc * CC_RS + cc_ph(16 * q + kk)], tw8[s]);
            dft8<false>(x);
#pragma unroll
            for (int s = 0; s < 8; s++) bufB[kk * CC_RS + cc_ph(8 * cc + s)] = x[s];
        }
        cluster.barrier_arrive();          // split barrier: this CTA is done reading its peers' A (waited on before A is refilled)
        arrived = true;
        __syncthreads();
        // ---- P4: row FFT over c -> |.| -> 10 log10 -> smoothing into y (:105-108)      B
        {
            double *y = pl.y + (size_t)ch * N;
            const bool bigchange = p.I[(size_t)I_ZERO_BB * p.cpad + ch] != 0;     // y[i]=20 pending (coarsefreqestimate.cpp:87)
            const int kk = threadIdx.x & 15, k2b = threadIdx.x >> 4;
            // y is only ever read by the fold search, for bins within [lo-epb-1, hi+epb+1] (coarsefreqestimate.cpp:112-130):
            // bins outside that window are neither loaded, evaluated (log10) nor stored
            const int need_lo = pl.lo - pl.expectedpeakbin - 1, need_hi = pl.hi + pl.expectedpeakbin + 1;
            double yo[8];                                                         // requested before the transform, consumed after it
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int i_sh = (16 * q + kk) + 128 * ((k2b + 16 * i + 64) & 127);
                yo[i] = (bigchange || i_sh < need_lo || i_sh > need_hi) ? 20.0 : y[i_sh];
            }
            cc_fft<false, false>(bufB, tws, [&](int f, int e, double2 v) { bufB[f * CC_RS + cc_ph(e)] = v; });
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int k2 = k2b + 16 * i;
                const int i_sh = (16 * q + kk) + 128 * ((k2 + 64) & 127);         // fftshift (:105)
                if (i_sh < need_lo || i_sh > need_hi) continue;
                const double2 x = bufB[kk * CC_RS + cc_ph(k2)];
                // 10*log10(max(|x|,1)) = 5*log10(max(|x|^2,1))
                y[i_sh] = yo[i] * 0.9 + 0.1 * 5 * log10_ge1(fmax(x.x * x.x + x.y * x.y, 1.0));   // :108
            }
        }
        __syncthreads();
    }
    if (arrived) cluster.barrier_wait();
    cluster.sync();                        // no CTA leaves while a peer may still read its shared memory
}

int cfe_cluster_run(const CfePlan &pl, const DemodParams &p, int oldest, int n_clusters, cudaStream_t s, long long *launches)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)(n_clusters * CC_CL)); cfg.blockDim = dim3(CC_T); cfg.dynamicSmemBytes = CC_SM_TOTAL; cfg.stream = s;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = CC_CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    JB_CUDA(cudaLaunchKernelEx(&cfg, cfe_cluster_kernel, pl, p, oldest));
    cfe_search_kernel<<<(p.n_channels + 3) / 4, 128, 0, s>>>(pl, p);
    JB_CUDA(cudaGetLastError());
    *launches += 2;
    return 0;
}
// number of clusters that can be co-resident (0: the device cannot run the cluster kernel)
int cfe_cluster_capacity()
{
    if (cudaFuncSetAttribute(cfe_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CC_SM_TOTAL) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(CC_CL * 64); cfg.blockDim = dim3(CC_T); cfg.dynamicSmemBytes = CC_SM_TOTAL;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = CC_CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, cfe_cluster_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (getenv("JAERO_DEBUG")) fprintf(stderr, "[jaero_b200] estimator clusters co-resident: %d x %d CTAs\n", n, CC_CL);
    return n;
}

__global__ void cfe_mark_kernel(int *flag, int value) { __threadfence(); atomicExch(flag, value); }
int cfe_mark_launch(int *flag, int value, cudaStream_t s)
{
    cfe_mark_kernel<<<1, 1, 0, s>>>(flag, value);
    JB_CUDA(cudaGetLastError());
    return 0;
}

// `oldest` = ring index of the oldest sample (the linearisation origin, oqpskdemodulator.cpp:418-424)
int cfe_run(const CfePlan &pl, const DemodParams &p, int oldest, cudaStream_t s, long long *launches)
{
    const int C = p.n_channels;
    for (int ch0 = 0; ch0 < C; ch0 += pl.group) {
        const int g = (C - ch0 < pl.group) ? C - ch0 : pl.group;
        dim3 gc(pl.n2 / TILE, g), gr(pl.n1 / TILE, g);
        cfe_col_kernel<false><<<gc, CFE_THREADS, 0, s>>>(pl, p.bb, pl.work_a, (size_t)p.bb_len, oldest, p.bb_len, ch0);
        cfe_row_mask_kernel<<<gr, CFE_THREADS, 0, s>>>(pl, pl.work_a, pl.work_b);
        cfe_col_kernel<true><<<gc, CFE_THREADS, 0, s>>>(pl, pl.work_b, pl.work_a, (size_t)pl.nfft, 0, pl.nfft, 0);
        cfe_row_logmag_kernel<<<gr, CFE_THREADS, 0, s>>>(pl, p, pl.work_a, ch0);
        *launches += 4;
    }
    cfe_search_kernel<<<(C + 3) / 4, 128, 0, s>>>(pl, p);
    *launches += 1;
    JB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace jb
