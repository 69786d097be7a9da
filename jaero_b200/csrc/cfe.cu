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
// Each pass moves 16-row / 16-column tiles (256 B segments) through shared memory; the radix-2
// butterflies of 16 independent n<=128-point FFTs run in one CTA. FFT rounding differs from the
// CPU oracle's radix-2 (different factorisation) at the 1e-13 level; the bin decision is integer.
#include "demod_device.cuh"

namespace jb {

static const int TILE = 16;
static const int MAXN = 128;
static const int CFE_THREADS = 256;

// in-place radix-2 DIT on TILE independent length-n sequences held bit-reversed in s[f][.]
// tw = W_N^k table, tw_stride = N/n.  inverse -> conjugated twiddles (unnormalised).
__device__ __forceinline__ void tile_fft(double2 (*s)[MAXN + 1], int n, int logn, const double2 *__restrict__ tw, int tw_stride, bool inverse)
{
    const int nb = n >> 1;                       // butterflies per sequence
    for (int st = 0; st < logn; st++) {
        const int half = 1 << st, len = half << 1;
        const int wstep = tw_stride * (n / len);
        for (int b = threadIdx.x; b < TILE * nb; b += CFE_THREADS) {
            const int f = b / nb, q = b - f * nb;
            const int grp = q >> st, j = q & (half - 1);
            const int i0 = grp * len + j, i1 = i0 + half;
            double2 w = tw[j * wstep];
            if (inverse) w.y = -w.y;
            const double2 x1 = s[f][i1], x0 = s[f][i0];
            const double2 t = make_double2(x1.x * w.x - x1.y * w.y, x1.x * w.y + x1.y * w.x);
            s[f][i0] = make_double2(x0.x + t.x, x0.y + t.y);
            s[f][i1] = make_double2(x0.x - t.x, x0.y - t.y);
        }
        __syncthreads();
    }
}
__device__ __forceinline__ int bitrev(int x, int bits) { return (int)(__brev((unsigned)x) >> (32 - bits)); }

// P1 / P3: column pass. grid = (n2/TILE, channels)
template <bool FUSED_INV_SQUARE>
__global__ void __launch_bounds__(CFE_THREADS)
cfe_col_kernel(CfePlan pl, const double2 *__restrict__ src, double2 *__restrict__ dst, size_t src_pitch, int rot, int ch0)
{
    __shared__ double2 s[TILE][MAXN + 1];
    const int ch = blockIdx.y;
    const int c0 = blockIdx.x * TILE;
    const int n1 = pl.n1, n2 = pl.n2, N = pl.nfft;
    const int l1 = 31 - __clz(n1);
    const double2 *in = src + (size_t)(ch0 + ch) * src_pitch;
    double2 *out = dst + (size_t)ch * N;
    // load column tile: element (r, c0+cc) of the n1 x n2 matrix, n = n2*r + c  (rot linearises the ring:
    // bbtmpbuff[j] = bbcycbuff[(ptr+j)%N], oqpskdemodulator.cpp:418-424)
    for (int e = threadIdx.x; e < n1 * TILE; e += CFE_THREADS) {
        const int r = e / TILE, cc = e - r * TILE;
        int n = n2 * r + c0 + cc;
        if (!FUSED_INV_SQUARE) { n += rot; if (n >= N) n -= N; }
        s[cc][bitrev(r, l1)] = in[n];
    }
    __syncthreads();
    if (FUSED_INV_SQUARE) {
        // column IFFT over k1 -> x'[n2*r+c]; square; then forward again
        tile_fft(s, n1, l1, pl.tw, N / n1, true);
        // square in natural order, then re-store bit-reversed for the forward transform
        double2 v[(MAXN * TILE) / CFE_THREADS];
        int cnt = 0;
        for (int e = threadIdx.x; e < n1 * TILE; e += CFE_THREADS, cnt++) {
            const int cc = e / n1, r = e - cc * n1;
            const double2 x = s[cc][r];
            v[cnt] = make_double2(x.x * x.x - x.y * x.y, x.x * x.y + x.y * x.x);   // in[i]*in[i] (:103)
        }
        __syncthreads();
        cnt = 0;
        for (int e = threadIdx.x; e < n1 * TILE; e += CFE_THREADS, cnt++) {
            const int cc = e / n1, r = e - cc * n1;
            s[cc][bitrev(r, l1)] = v[cnt];
        }
        __syncthreads();
    }
    tile_fft(s, n1, l1, pl.tw, N / n1, false);
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
    const int l2 = 31 - __clz(n2);
    const double2 *in = src + (size_t)ch * N;
    double2 *out = dst + (size_t)ch * N;
    for (int e = threadIdx.x; e < n2 * TILE; e += CFE_THREADS) {
        const int rr = e / n2, c = e - rr * n2;
        s[rr][bitrev(c, l2)] = in[(size_t)(r0 + rr) * n2 + c];
    }
    __syncthreads();
    tile_fft(s, n2, l2, pl.tw, N / n2, false);
    // X[k1 + n1*k2] sits at s[k1-r0][k2]; mask (:99-100), then re-store bit-reversed for the inverse
    double2 v[(MAXN * TILE) / CFE_THREADS];
    int cnt = 0;
    for (int e = threadIdx.x; e < n2 * TILE; e += CFE_THREADS, cnt++) {
        const int rr = e / n2, k2 = e - rr * n2;
        const int k = (r0 + rr) + n1 * k2;
        double2 x = s[rr][k2];
        if (!pl.is8400) { if (k >= pl.startbin && k <= pl.stopbin) x = make_double2(0.0, 0.0); }
        else { const double w = pl.window[k]; x = make_double2(x.x * w, x.y * w); }
        v[cnt] = x;
    }
    __syncthreads();
    cnt = 0;
    for (int e = threadIdx.x; e < n2 * TILE; e += CFE_THREADS, cnt++) {
        const int rr = e / n2, k2 = e - rr * n2;
        s[rr][bitrev(k2, l2)] = v[cnt];
    }
    __syncthreads();
    tile_fft(s, n2, l2, pl.tw, N / n2, true);
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
    const int l2 = 31 - __clz(n2);
    const double2 *in = src + (size_t)ch * N;
    double *y = pl.y + (size_t)(ch0 + ch) * N;
    const bool bigchange = p.I[(size_t)I_ZERO_BB * p.cpad + ch0 + ch] != 0;     // y[i]=20 pending (coarsefreqestimate.cpp:87)
    for (int e = threadIdx.x; e < n2 * TILE; e += CFE_THREADS) {
        const int rr = e / n2, c = e - rr * n2;
        s[rr][bitrev(c, l2)] = in[(size_t)(r0 + rr) * n2 + c];
    }
    __syncthreads();
    tile_fft(s, n2, l2, pl.tw, N / n2, false);
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
    for (int i = pl.lo + lane; i < pl.hi; i += 32) {
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

int cfe_run(const CfePlan &pl, const DemodParams &p, int bb_pos, cudaStream_t s, long long *launches)
{
    const int C = p.n_channels;
    for (int ch0 = 0; ch0 < C; ch0 += pl.group) {
        const int g = (C - ch0 < pl.group) ? C - ch0 : pl.group;
        dim3 gc(pl.n2 / TILE, g), gr(pl.n1 / TILE, g);
        cfe_col_kernel<false><<<gc, CFE_THREADS, 0, s>>>(pl, p.bb, pl.work_a, (size_t)pl.nfft, bb_pos, ch0);
        cfe_row_mask_kernel<<<gr, CFE_THREADS, 0, s>>>(pl, pl.work_a, pl.work_b);
        cfe_col_kernel<true><<<gc, CFE_THREADS, 0, s>>>(pl, pl.work_b, pl.work_a, (size_t)pl.nfft, 0, 0);
        cfe_row_logmag_kernel<<<gr, CFE_THREADS, 0, s>>>(pl, p, pl.work_a, ch0);
        *launches += 4;
    }
    cfe_search_kernel<<<(C + 3) / 4, 128, 0, s>>>(pl, p);
    *launches += 1;
    JB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace jb
