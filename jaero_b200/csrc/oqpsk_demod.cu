// K1a — fused continuous OQPSK demodulator segment kernel (8400* / 10500 bps), one thread per channel.
//   (* the 8400 bps FFT pre-filter, oqpskdemodulator.cpp:343-381, is not part of this kernel yet)
//
// Replaces OqpskDemodulator::writeData (JAERO/oqpskdemodulator.cpp:334-627) and
// OqpskDemodulator::FreqOffsetEstimateSlot (:629-677) for a whole batch of channels:
// int16 -> coarse-estimator ring write -> NCO mix (table NCO, DSP.cpp:79-85) -> 55-tap RRC FIR x2
// (DSP.cpp:292-304) -> EbNo (DSP.cpp:729-744) -> AGC (DSP.cpp:370-379) -> clip -> symbol-timing chain
// (delays, resonator, T/8 quadrature, arg, PLL nudges :473-484) -> strobe interpolation (:488-494)
// -> carrier loop (:512-532) -> bias rotate / 400-symbol delay (:535-537) -> MSE gate (:563-565)
// -> soft bits (:569-592).
//
// The per-sample recursion is serial per channel (the carrier loop feeds the NCO ahead of the
// FIR), so the parallel axis is the channel: lane = channel, all sample-rate ring positions are
// warp-uniform, every HBM ring access is a coalesced 256 B row. FIR delay lines live in shared
// memory ([tap][lane], conflict-free), everything else in registers.
#include "demod_device.cuh"

namespace jb {

static const int OQ_THREADS = 32;
static const int OQ_NT1 = 56;             // 55 taps + 1 (FIR ring, DSP.cpp:277)
static const int OQ_FIRROWS = 2 * OQ_NT1; // every entry is stored twice so any 55-entry window is contiguous
static const int OQ_EBNO_TAIL = 256;      // samples before the end of a launch over which the EbNo read-out is evaluated
static const int OQ_T = 32;               // tile length (samples) of the staged HBM streams; one 256 B ring row per lane
static const int OQ_PROW = 80;            // bytes per channel row of a PCM tile in shared memory (64 B payload + pad)
// shared memory map (bytes)
static const int OQ_SM_FIR = 2 * OQ_FIRROWS * OQ_THREADS * 8;          // FIR windows re/im
static const int OQ_SM_RING = OQ_T * OQ_THREADS * 8;                   // one ring tile [T][32] doubles
static const int OQ_SM_PCM = OQ_THREADS * OQ_PROW;                     // one PCM tile
static const int OQ_XROW = 32 * 16 + 16;    // bytes per channel row of a pre-filtered-sample tile (32 double2 + pad)
static const int OQ_SM_X = OQ_THREADS * OQ_XROW;
static const int OQ_SM_TOTAL = OQ_SM_FIR + 6 * OQ_SM_RING + 2 * OQ_SM_PCM + 64;
static const int OQ_SM_TOTAL_PRE = OQ_SM_TOTAL + 2 * OQ_SM_X;

#define LD(idx) p.D[(size_t)(idx) * cpad + ch]
#define LI(idx) p.I[(size_t)(idx) * cpad + ch]

// 55-tap FIR over a contiguous window (oldest first), exactly the accumulation order of
// FIR::FIRUpdateAndProcess (DSP.cpp:296-303): outsum += points[i]*buff[tptr], i = 0..54.
__device__ __forceinline__ void fir55(const DemodParams &p, const double *__restrict__ wre, const double *__restrict__ wim, double &ore, double &oim)
{
    double sre = 0, sim = 0;
#pragma unroll
    for (int k = 0; k < 55; k++) {
        sre += p.taps[k] * wre[k * OQ_THREADS];
        sim += p.taps[k] * wim[k * OQ_THREADS];
    }
    ore = sre; oim = sim;
}
// the first 54 terms of the same sum (everything except the newest sample, which is still being mixed)
__device__ __forceinline__ void fir54(const DemodParams &p, const double *__restrict__ wre, const double *__restrict__ wim, double &ore, double &oim)
{
    double sre = 0, sim = 0;
#pragma unroll
    for (int k = 0; k < 54; k++) {
        sre += p.taps[k] * wre[k * OQ_THREADS];
        sim += p.taps[k] * wim[k * OQ_THREADS];
    }
    ore = sre; oim = sim;
}

// PRE = true: the 8400 bps variant (oqpskdemodulator.cpp:436-448): no FIR in the loop, the sample entering the loop is
// mixer2.CIS * cval_prefiltered[i] (K6 output, staged like the PCM rows), and mixer2's frequency is summed per sample.
template <bool PRE>
__global__ void __launch_bounds__(OQ_THREADS)
oqpsk_segment_kernel(const __grid_constant__ DemodParams p, const SegmentArgs a, const int16_t *__restrict__ pcm, size_t stride,
                     const double2 *__restrict__ xpre, size_t xstride, double *__restrict__ m2_freq_sum)
{
    extern __shared__ __align__(128) unsigned char oq_smem_raw[];
    double *s_re = reinterpret_cast<double *>(oq_smem_raw);   // [OQ_FIRROWS][32]
    double *s_im = s_re + OQ_FIRROWS * OQ_THREADS;
    double *t_agc = reinterpret_cast<double *>(oq_smem_raw + OQ_SM_FIR);          // [2][T][32]
    double *t_e1 = t_agc + 2 * OQ_T * OQ_THREADS;
    double *t_e2 = t_e1 + 2 * OQ_T * OQ_THREADS;
    unsigned char *t_pcm = oq_smem_raw + OQ_SM_FIR + 6 * OQ_SM_RING;              // [2][32][OQ_PROW]
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(t_pcm + 2 * OQ_SM_PCM);   // ring[2], pcm[2], x[2]
    unsigned char *t_x = oq_smem_raw + OQ_SM_TOTAL;                               // [2][32][OQ_XROW] (PRE only)
    const int lane = threadIdx.x;
    const int ch_raw = blockIdx.x * OQ_THREADS + lane;
    const bool live = ch_raw < p.n_channels;
    const int ch = ch_raw;                                    // dead lanes run on their (allocated) pad column with zero input
    const int nlive = min(OQ_THREADS, p.n_channels - (int)blockIdx.x * OQ_THREADS);
    const size_t cpad = p.cpad;
    if (lane == 0) { for (int k = 0; k < 6; k++) mbar_init(&bars[k], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    // ---------------- load state
    Osc m2 = {LD(D_M2_PTR), LD(D_M2_STEP), LD(D_M2_FREQ), LD(D_M2_LAST)};
    Osc mc = {LD(D_MC_PTR), LD(D_MC_STEP), LD(D_MC_FREQ), LD(D_MC_LAST)};
    Osc st = {LD(D_ST_PTR), LD(D_ST_STEP), LD(D_ST_FREQ), LD(D_ST_LAST)};
    Osc sr = {LD(D_SR_PTR), LD(D_SR_STEP), LD(D_SR_FREQ), LD(D_SR_LAST)};
    double agc_sum = LD(D_AGC_SUM), agc_val = LD(D_AGC_VAL);
    double eb_sum1 = LD(D_EB_SUM1), eb_sum2 = LD(D_EB_SUM2), eb_ebno = LD(D_EB_EBNO);
    double dly_s0 = LD(D_DLY_S0);
    double d41_0 = LD(D_DLY41_0), d41_1 = LD(D_DLY41_1), d41_2 = LD(D_DLY41_2);
    double d42_0 = LD(D_DLY42_0), d42_1 = LD(D_DLY42_1), d42_2 = LD(D_DLY42_2);
    double d8_0 = LD(D_DLY8_0), d8_1 = LD(D_DLY8_1), d8_2 = LD(D_DLY8_2);
    Biquad res = {LD(D_RES_X1), LD(D_RES_X2), LD(D_RES_Y1), LD(D_RES_Y2)};
    Biquad lf = {LD(D_LF_X1), LD(D_LF_X2), LD(D_LF_Y1), LD(D_LF_Y2)};
    double2 sig2_last = make_double2(LD(D_SIG2L_RE), LD(D_SIG2L_IM));
    double2 pt_d = make_double2(LD(D_PTD_RE), LD(D_PTD_IM));
    double marg_sum = LD(D_MARG_SUM), marg_val = LD(D_MARG_VAL);
    double pm_sum = LD(D_MSE_PM_SUM), ma_sum = LD(D_MSE_MA_SUM), mse = LD(D_MSE);
    double lastmse = LD(D_LASTMSE);
    double2 sc0 = make_double2(LD(D_SCAT0_RE), LD(D_SCAT0_IM)), sc1 = make_double2(LD(D_SCAT1_RE), LD(D_SCAT1_IM));
    int yui = LI(I_YUI), countdown = LI(I_COUNTDOWN), countdown2 = LI(I_COUNTDOWN2), dcd = LI(I_DCD);
    int sig2l_init = LI(I_SIG2L_INIT);
    int marg_pos = LI(I_MARG_POS), dt_pos = LI(I_DT_POS), mse_pos = LI(I_MSE_POS);
    int soft_count = LI(I_SOFT_COUNT), soft_pending = LI(I_SOFT_PENDING), soft_overflow = LI(I_SOFT_OVERFLOW);
    int sig_true = LI(I_SIG_TRUE), sig_false = LI(I_SIG_FALSE);

    for (int k = 0; k < OQ_NT1; k++) {
        const double vr = p.fir_re[(size_t)k * cpad + ch], vi = p.fir_im[(size_t)k * cpad + ch];
        s_re[k * OQ_THREADS + lane] = vr; s_re[(k + OQ_NT1) * OQ_THREADS + lane] = vr;
        s_im[k * OQ_THREADS + lane] = vi; s_im[(k + OQ_NT1) * OQ_THREADS + lane] = vi;
    }
    if (a.new_write) lastmse = mse;                                       // oqpskdemodulator.cpp:339

    // ---------------- FreqOffsetEstimateSlot (oqpskdemodulator.cpp:629-677), re-entrant in the reference:
    // it runs after the ring write and before the mixer of the same sample.
    if (a.apply_cfe) {
        const double est = p.cfe_est_out[ch];
        if ((mse < p.signalthreshold) && (!dcd)) {                        // :642-650
            if (countdown2 > 0) countdown2--;
            else osc_set_freq(m2, mc.freq + est, p.Fs);
        } else countdown2 = 5;
        if ((mse > p.signalthreshold) && (fabs(m2.freq - (mc.freq + est)) > 3.0))    // :653-657
            osc_set_freq(m2, mc.freq + est, p.Fs);
        if ((p.afc) && (mse < p.signalthreshold) && (fabs(m2.freq - mc.freq) > 3.0)) {   // :658-669
            if (countdown > 0) countdown--;
            else {
                osc_set_freq(mc, m2.freq, p.Fs);
                if (mc.freq < p.lockingbw / 2.0) osc_set_freq(mc, p.lockingbw / 2.0, p.Fs);
                if (mc.freq > (p.Fs / 2.0 - p.lockingbw / 2.0)) osc_set_freq(mc, p.Fs / 2.0 - p.lockingbw / 2.0, p.Fs);
                LI(I_EMPTYING) = 4;                                       // CoarseFreqEstimate::bigchange (coarsefreqestimate.cpp:84-88)
                LI(I_ZERO_BB) = 1;                                        // y[]=20 is applied by the estimator kernel on its next run
                double2 *rowz = p.bb + (size_t)ch * p.bb_len;             // :667 bbcycbuff[j]=0
                if (live) for (int j = 0; j < p.bb_len; j++) rowz[j] = make_double2(0.0, 0.0);
            }
        } else countdown = 4;
        if (mse > p.signalthreshold) { sig_false++; if (p.wire_sigstat) { { const int ln_ = LI(I_LOST_N); if (ln_ < LOST_CAP) p.lost_pos[(size_t)ln_ * cpad + ch] = LI(I_SOFT_COUNT); LI(I_LOST_N) = ln_ + 1; LI(I_DCD) = 0; } dcd = 0; } } else sig_true++;       // :674-675
    }

    // ---------------- lock-step positions
    const int agc_len = p.agc_len, eb_len = p.ebno_len;
    long long S = a.sample0;                                  // samples fully processed so far: drives every sample-rate ring
    int fir_pos = (int)(S % OQ_NT1);                          // next FIR slot to write
    int bb_pos = a.bb_pos, coarse_counter = a.coarse_counter;
    const bool ebno_on = p.report_ebno != 0;
    const int16_t *row = pcm + (size_t)ch * stride;
    double2 *bb_row = p.bb + (size_t)ch * p.bb_len;
    const int bbn = p.bb_len;
    const bool cpu_reduce = p.cpu_reduce != 0;
    const double Fs = p.Fs, fbr = p.fb, thr = p.signalthreshold, ee = p.ee;
    const double res_a1 = p.res_a1, res_a2 = p.res_a2, res_b0 = p.res_b0, res_b1 = p.res_b1, res_b2 = p.res_b2;
    int p41 = (int)(S % (p.k41 + 1)), p8 = (int)(S % (p.k8 + 1));   // Delay<> ring positions (lock-step)
    const int k41 = p.k41, k8 = p.k8;
    const double *__restrict__ cos_t = p.cos_t, *__restrict__ sin_t = p.sin_t;
    const int marg_len = p.marg_len, dt_len = p.dt_len, mse_len = p.mse_len;
    const int eb_from = a.i1 - OQ_EBNO_TAIL;
    const long long S_end = S + (a.i1 - a.i0) - (a.stop_after_a ? 1 : 0);   // value of S when this launch returns

    // ---------------- HBM streams staged through shared memory by the bulk-copy (TMA) engine.
    // The three sample-rate rings (AGC, EbNo E and E2: [slot][channel], 256 B per slot for this warp's 32 channels) and the
    // PCM rows are moved in tiles of 32 samples: lane r copies ring row r / its own PCM row with cp.async.bulk, completion
    // is signalled on an mbarrier, two buffers per stream, the next tile is in flight while the current one is consumed and
    // updated in place; finished ring tiles go back to HBM with bulk stores. The per-sample loop therefore never waits on
    // DRAM: its ring/PCM operands are shared-memory reads.
    auto ring_rows = [&](long long tile, double *&g_agc, double *&g_e1, double *&g_e2) {
        const long long s0 = tile * OQ_T;
        g_agc = p.agc_ring + ((size_t)(s0 % agc_len) + lane) * cpad + (size_t)blockIdx.x * OQ_THREADS;
        if (ebno_on) {
            g_e1 = p.ebno_e1 + ((size_t)(s0 % eb_len) + lane) * cpad + (size_t)blockIdx.x * OQ_THREADS;
            g_e2 = p.ebno_e2 + ((size_t)(s0 % eb_len) + lane) * cpad + (size_t)blockIdx.x * OQ_THREADS;
        }
    };
    const unsigned ring_tx = (ebno_on ? 3u : 1u) * OQ_SM_RING;
    auto ring_load = [&](long long tile) {                    // all lanes call; lane r moves row r of the tile
        const int b = (int)(tile & 1);
        fence_proxy_async();
        if (lane == 0) mbar_expect_tx(&bars[b], ring_tx);
        __syncwarp();
        double *g_agc = nullptr, *g_e1 = nullptr, *g_e2 = nullptr;
        ring_rows(tile, g_agc, g_e1, g_e2);
        bulk_g2s(t_agc + (b * OQ_T + lane) * OQ_THREADS, g_agc, OQ_THREADS * 8, &bars[b]);
        if (ebno_on) {
            bulk_g2s(t_e1 + (b * OQ_T + lane) * OQ_THREADS, g_e1, OQ_THREADS * 8, &bars[b]);
            bulk_g2s(t_e2 + (b * OQ_T + lane) * OQ_THREADS, g_e2, OQ_THREADS * 8, &bars[b]);
        }
    };
    auto ring_store = [&](long long tile) {                   // write the (in-place updated) tile back to HBM
        const int b = (int)(tile & 1);
        fence_proxy_async();
        __syncwarp();
        double *g_agc = nullptr, *g_e1 = nullptr, *g_e2 = nullptr;
        ring_rows(tile, g_agc, g_e1, g_e2);
        bulk_s2g(g_agc, t_agc + (b * OQ_T + lane) * OQ_THREADS, OQ_THREADS * 8);
        if (ebno_on) {
            bulk_s2g(g_e1, t_e1 + (b * OQ_T + lane) * OQ_THREADS, OQ_THREADS * 8);
            bulk_s2g(g_e2, t_e2 + (b * OQ_T + lane) * OQ_THREADS, OQ_THREADS * 8);
        }
        bulk_commit();
    };
    // PCM: tile t covers buffer samples [32t, 32t+32) of every channel row (16 B aligned: stride % 8 == 0, host-checked)
    auto pcm_bytes = [&](int tile) -> unsigned {
        long long left = (long long)stride - (long long)tile * OQ_T;
        if (left > OQ_T) left = OQ_T;
        return left > 0 ? (unsigned)(left * 2) : 0u;
    };
    auto pcm_load = [&](int tile) {
        const int b = tile & 1;
        const unsigned nb = pcm_bytes(tile);
        fence_proxy_async();
        if (lane == 0) mbar_expect_tx(&bars[2 + b], nb * (unsigned)nlive);
        __syncwarp();
        if (live && nb) bulk_g2s(t_pcm + b * OQ_SM_PCM + lane * OQ_PROW, row + (size_t)tile * OQ_T, nb, &bars[2 + b]);
        if (PRE) {
            long long left = (long long)xstride - (long long)tile * OQ_T;
            if (left > OQ_T) left = OQ_T;
            const unsigned xb = left > 0 ? (unsigned)(left * 16) : 0u;
            if (lane == 0) mbar_expect_tx(&bars[4 + b], xb * (unsigned)nlive);
            __syncwarp();
            if (live && xb) bulk_g2s(t_x + b * OQ_SM_X + lane * OQ_XROW, xpre + (size_t)ch * xstride + (size_t)tile * OQ_T, xb, &bars[4 + b]);
        }
    };
    unsigned phases = 0u;                                     // expected parity per barrier (bit b: ring b, bit 2+b: pcm b)
#define OQ_WAIT(idx) do { mbar_wait(&bars[(idx)], (phases >> (idx)) & 1u); phases ^= (1u << (idx)); } while (0)
    long long rt = S / OQ_T;                                  // current ring tile
    int pt = a.i0 / OQ_T;                                     // current PCM tile
    bool ring_next_issued = false, pcm_next_issued = false;
    ring_load(rt);
    pcm_load(pt);
    if ((rt + 1) * OQ_T < S_end) { ring_load(rt + 1); ring_next_issued = true; }
    if ((pt + 1) * OQ_T < a.i1) { pcm_load(pt + 1); pcm_next_issued = true; }
    OQ_WAIT((int)(rt & 1));
    OQ_WAIT(2 + (pt & 1));
    if (PRE) OQ_WAIT(4 + (pt & 1));
    bool ring_dirty = false;

    // ---- table / symbol-rate operands requested ahead of use (the SM issues in order: a load stalls the warp only when
    // its result is consumed)
    double c2_re, c2_im, cs_re, cs_im, cc_re, cc_im;
    { const int t = osc_index(m2.ptr); c2_re = cos_t[t]; c2_im = sin_t[t]; }
    { const int t = osc_index(st.ptr); cs_re = cos_t[t]; cs_im = sin_t[t]; }
    { const int t = osc_index(mc.ptr); cc_re = cos_t[t]; cc_im = sin_t[t]; }
    double sy_marg_old = p.marg_ring[(size_t)marg_pos * cpad + ch];
    double sy_pm_old = p.mse_pm[(size_t)mse_pos * cpad + ch];
    double sy_ma_old = p.mse_ma[(size_t)mse_pos * cpad + ch];
    double2 sy_dt_old;
    { int r = dt_pos + 1; if (r >= dt_len) r = 0; sy_dt_old = p.dt_ring[(size_t)r * cpad + ch]; }
    int4 pk = make_int4(0, 0, 0, 0);                          // 8 consecutive PCM samples of this lane's channel
    bool pk_valid = false;

    // FIR output of the first sample of this launch: the 55 entries older than the slot about to be written
    // (DSP.cpp:292-304: the output excludes the sample just stored). The window that ends at logical slot q starts at
    // row q+2 of the doubled buffer.
    double fre = 0, fim = 0;
    if (!PRE) {
        int newest = fir_pos - 1; if (newest < 0) newest += OQ_NT1;
        fir55(p, s_re + (newest + 2) * OQ_THREADS + lane, s_im + (newest + 2) * OQ_THREADS + lane, fre, fim);
    }
    double m2sum = PRE ? (a.new_write ? 0.0 : m2_freq_sum[ch]) : 0.0;     // mixer2_freq_sum (:385,447)

    for (int i = a.i0; i < a.i1; i++) {
        // ---- PCM sample from the staged tile
        const int po = i & (OQ_T - 1);
        if ((i >> 5) != pt) {                                 // entered the next PCM tile (warp-uniform)
            pt = i >> 5;
            OQ_WAIT(2 + (pt & 1));
            if (PRE) OQ_WAIT(4 + (pt & 1));
            pcm_next_issued = false;
            if ((pt + 1) * OQ_T < a.i1) { pcm_load(pt + 1); pcm_next_issued = true; }
            pk_valid = false;
        }
        if (!pk_valid || (po & 7) == 0) {
            pk = *reinterpret_cast<const int4 *>(t_pcm + (pt & 1) * OQ_SM_PCM + lane * OQ_PROW + (po >> 3) * 16);
            pk_valid = true;
        }
        int cur_pcm;
        {
            const int k = po & 7;
            const int w = (k < 2) ? pk.x : (k < 4) ? pk.y : (k < 6) ? pk.z : pk.w;
            cur_pcm = (k & 1) ? (w >> 16) : (int)(short)(w & 0xffff);
            if (!live) cur_pcm = 0;
        }
        const double dval = ((double)cur_pcm) / 32768.0;                  // :390

        // ---- A: coarse-estimator ring (:410-429); the host ends the segment on the trigger sample
        if (!(i == a.i0 && a.skip_a_first)) {
            if (coarse_counter >= Fs || !cpu_reduce) {
                if (live) bb_row[bb_pos] = make_double2(cc_re * dval, cc_im * dval);
                bb_pos++; if (bb_pos >= bbn) bb_pos = 0;
            }
        }
        if (i == a.i1 - 1 && a.stop_after_a) break;
        coarse_counter++;                                                 // :431
        // mixer_center only free-runs inside the loop: advance it now (:601) and request its next table entry a whole
        // iteration before the ring write that consumes it
        osc_next_frame(mc);
        { const int t = osc_index(mc.ptr); cc_re = cos_t[t]; cc_im = sin_t[t]; }
        // speculative request for mixer2's next entry (right unless this sample turns out to be a carrier-update strobe)
        const int m2_spec = osc_next_index(m2);
        double n2_re = cos_t[m2_spec], n2_im = sin_t[m2_spec];

        // ---- B. cval = CIS * dval (:453) goes into the FIR ring; the output for THIS sample (fre,fim) was formed from
        // the older entries in the previous iteration, the output for the NEXT sample is formed now — an independent
        // dependency chain the scheduler interleaves with the serial loop arithmetic below.
        // The 54 older terms are summed first (same order as DSP.cpp:296-303), the newest term is appended once the mixed
        // sample is available.
        double nfre = 0, nfim = 0;
        if (!PRE) fir54(p, s_re + (fir_pos + 2) * OQ_THREADS + lane, s_im + (fir_pos + 2) * OQ_THREADS + lane, nfre, nfim);
        if (PRE) {
            // sig2 = mixer2.WTCISValue()*cval_prefiltered[i] (:440); mixer2_freq_sum+=mixer2.GetFreqHz() (:447)
            double2 xv = *reinterpret_cast<const double2 *>(t_x + (pt & 1) * OQ_SM_X + lane * OQ_XROW + po * 16);
            if (!live) xv = make_double2(0.0, 0.0);
            const double2 sg = cmul(make_double2(c2_re, c2_im), xv);
            fre = sg.x; fim = sg.y;
            m2sum += m2.freq;
        }

        const double sre = fre, sim = fim;
        const double dabval = sqrt(sre * sre + sim * sim);                // :461

        const int ro = (int)(S & (OQ_T - 1));
        const int rslot = (((int)(rt & 1)) * OQ_T + ro) * OQ_THREADS + lane;   // this sample's slot in the staged ring tiles
        if (ebno_on) {                                                    // OQPSKEbNoMeasure::Update (DSP.cpp:729-744)
            const double sq = dabval * dabval;
            eb_sum2 = eb_sum2 - t_e2[rslot]; eb_sum2 = eb_sum2 + fabs(sq); t_e2[rslot] = fabs(sq);
            eb_sum1 = eb_sum1 - t_e1[rslot]; eb_sum1 = eb_sum1 + fabs(dabval); t_e1[rslot] = fabs(dabval);
            // The smoothed read-out EbNo <- 0.8 EbNo + 0.2 tebno forgets its past by 0.8^k: evaluating it over the last
            // 256 samples of a launch reproduces the value a per-sample evaluation has at the end of the launch to
            // below 1e-24 relative, without a log10 and three divisions on every sample. Observable only (DSP.h:250).
            if (i >= eb_from) {
                const double e2val = eb_sum2 / ((double)eb_len), mean = eb_sum1 / ((double)eb_len);
                const double mean_sq = mean * mean;
                double var = (e2val) - (mean * mean);
                var -= (0.024709 * mean_sq);
                double mvr = (((Fs * mean_sq / (2.0 * fbr * var))) * 0.13743);
                if (mvr < 0.000000001) mvr = 0.000000001;
                double tebno = 10.0 * log10(mvr);
                if (isnan(tebno)) tebno = 50;
                if (tebno > 50.0) tebno = 50;
                if (tebno < 0.0) tebno = 0;
                eb_ebno = eb_ebno * 0.8 + 0.2 * tebno;
            }
        }

        {   // AGC::Update (DSP.cpp:370-379)
            agc_sum = agc_sum - t_agc[rslot];
            agc_sum = agc_sum + fabs(dabval);
            t_agc[rslot] = fabs(dabval);
            ring_dirty = true;
            agc_val = 1.414213562 / fmax(agc_sum / ((double)agc_len), 0.000001);
            agc_val = fmax(agc_val, 0.000001);
        }
        double2 sig2 = make_double2(sre * agc_val, sim * agc_val);        // :466
        const double abval = hypot(sig2.x, sig2.y);                       // :469 std::abs
        if (abval > 2.84) { const double g = (2.84 / abval); sig2 = make_double2(g * sig2.x, g * sig2.y); }   // :470

        // ---- symbol timing (:473-484)
        const double ab2 = abval * abval;
        const double st_diff = (0.0 * ab2 + (1.0 - 0.0) * dly_s0) - (ab2);    // Delay(1): weighting 0 -> x[n-1]
        dly_s0 = ab2;
        // Delay(T/4): older = x[n-k41], newer = x[n-k41+1]  (DSP.h:357-374)
        double st_d1out, st_d2out;
        const double w41 = p.w41v[p41], w8 = p.w8v[p8];
        p41++; if (p41 > k41) p41 = 0;
        p8++; if (p8 > k8) p8 = 0;
        {
            const double older = (k41 == 3) ? d41_2 : (k41 == 2 ? d41_1 : d41_0);
            const double newer = (k41 == 3) ? d41_1 : (k41 == 2 ? d41_0 : st_diff);
            st_d1out = (w41 * newer + (1.0 - w41) * older);
            d41_2 = d41_1; d41_1 = d41_0; d41_0 = st_diff;
        }
        {
            const double older = (k41 == 3) ? d42_2 : (k41 == 2 ? d42_1 : d42_0);
            const double newer = (k41 == 3) ? d42_1 : (k41 == 2 ? d42_0 : st_d1out);
            st_d2out = (w41 * newer + (1.0 - w41) * older);
            d42_2 = d42_1; d42_1 = d42_0; d42_0 = st_d1out;
        }
        double st_eta = (st_d2out - st_diff) * st_d1out;
        st_eta = biquad_update(res, st_eta, res_a1, res_a2, res_b0, res_b1, res_b2);
        double d8out;
        {
            const double older = (k8 == 3) ? d8_2 : (k8 == 2 ? d8_1 : d8_0);
            const double newer = (k8 == 3) ? d8_1 : (k8 == 2 ? d8_0 : st_eta);
            d8out = (w8 * newer + (1.0 - w8) * older);
            d8_2 = d8_1; d8_1 = d8_0; d8_0 = st_eta;
        }
        const double2 st_out = cmul(make_double2(cs_re, cs_im), make_double2(st_eta, -d8out));   // :478-479
        const double st_angle_error = atan2_fast(st_out.y, st_out.x);          // :480 std::arg
        osc_set_freq(st, (-st_angle_error * 0.00000001) + st.freq, Fs);   // :481 IncreseFreqHz
        osc_advance_fraction_of_wave(st, -st_angle_error * 0.01 / 360.0); // :482
        if (st.freq < (sr.freq - 0.1)) osc_set_freq(st, (sr.freq - 0.1), Fs);
        if (st.freq > (sr.freq + 0.1)) osc_set_freq(st, (sr.freq + 0.1), Fs);

        if (!sig2l_init) { sig2_last = sig2; sig2l_init = 1; }            // :487 static initialiser
        double frac;
        if (osc_have_passed_point(st, ee, frac)) {                        // :488
            const double pt_last = frac, pt_this = 1.0 - pt_last;
            const double2 pt = make_double2(pt_this * sig2.x + pt_last * sig2_last.x, pt_this * sig2.y + pt_last * sig2_last.y);
            yui ^= 1;                                                     // yui++; yui%=2;
            if (!yui) pt_d = pt;
            else {
                double2 pt_qpsk = make_double2(pt.x, pt_d.y);             // :503
                const double ct_xt = tanh(pt.y) * pt.x;
                const double ct_xt_d = tanh(pt_d.x) * pt_d.y;
                double ct_ec = ct_xt_d - ct_xt;
                if (ct_ec > M_PI) ct_ec = M_PI;
                if (ct_ec < -M_PI) ct_ec = -M_PI;
                if (fbr > 8400) {                                         // :518-525
                    ct_ec = biquad_update(lf, ct_ec, p.lf_a1, p.lf_a2, p.lf_b0, p.lf_b1, p.lf_b2);
                    if (ct_ec > M_PI_2) ct_ec = M_PI_2;
                    if (ct_ec < -M_PI_2) ct_ec = -M_PI_2;
                    osc_increase_phase_deg(m2, 1.0 * ct_ec);
                    osc_set_freq(m2, (0.01 * ct_ec) + m2.freq, Fs);
                } else {                                                  // :526-532
                    osc_increase_phase_deg(m2, 1.0 * ct_ec);
                    const double lfo = biquad_update(lf, ct_ec, p.lf_a1, p.lf_a2, p.lf_b0, p.lf_b1, p.lf_b2);
                    osc_set_freq(m2, (0.5 * 0.01 * lfo) + m2.freq, Fs);
                }
                {   // marg->UpdateSigned(ct_ec)  MA(800)  (:535, DSP.cpp:418-426)
                    marg_sum = marg_sum - sy_marg_old;
                    marg_sum = marg_sum + (ct_ec);
                    p.marg_ring[(size_t)marg_pos * cpad + ch] = (ct_ec);
                    marg_pos++; if (marg_pos >= marg_len) marg_pos = 0;
                    marg_val = marg_sum / ((double)marg_len);
                }
                {   // dt.update(pt_qpsk): 400-symbol delay (:536, DSP.h:455-460)
                    p.dt_ring[(size_t)dt_pos * cpad + ch] = pt_qpsk;
                    dt_pos++; if (dt_pos >= dt_len) dt_pos = 0;
                    pt_qpsk = sy_dt_old;                                  // requested after the previous strobe
                }
                pt_qpsk = cmul(pt_qpsk, make_double2(cos(marg_val), sin(marg_val)));   // :537
                sc1 = sc0; sc0 = pt_qpsk;
                {   // MSEcalc::Update (DSP.cpp:451-463)
                    const size_t e = (size_t)mse_pos * cpad + ch;
                    const double ab = hypot(pt_qpsk.x, pt_qpsk.y);
                    pm_sum = pm_sum - sy_pm_old; pm_sum = pm_sum + fabs(ab); p.mse_pm[e] = fabs(ab);
                    double mu = pm_sum / ((double)mse_len);
                    if (mu < 0.000001) mu = 0.000001;
                    const double r2 = sqrt(2.0);
                    const double tre = (r2 * pt_qpsk.x) / mu, tim = (r2 * pt_qpsk.y) / mu;
                    const double tda = (fabs(tre) - 1.0), tdb = (fabs(tim) - 1.0);
                    const double v = (tda * tda) + (tdb * tdb);
                    ma_sum = ma_sum - sy_ma_old; ma_sum = ma_sum + fabs(v); p.mse_ma[e] = fabs(v);
                    mse_pos++; if (mse_pos >= mse_len) mse_pos = 0;
                    mse = ma_sum / ((double)mse_len);
                }
                // operands of the next strobe pair (slots written >= 400 symbols ago)
                sy_marg_old = p.marg_ring[(size_t)marg_pos * cpad + ch];
                sy_pm_old = p.mse_pm[(size_t)mse_pos * cpad + ch];
                sy_ma_old = p.mse_ma[(size_t)mse_pos * cpad + ch];
                { int r = dt_pos + 1; if (r >= dt_len) r = 0; sy_dt_old = p.dt_ring[(size_t)r * cpad + ch]; }
                if (live && mse < thr) {                                  // :565
                    push_soft(p, ch, soft_count, soft_pending, soft_overflow, q_round(0.75 * pt_qpsk.y * 127.0 + 128.0));
                    push_soft(p, ch, soft_count, soft_pending, soft_overflow, q_round(0.75 * pt_qpsk.x * 127.0 + 128.0));
                    if (soft_pending >= 32) {                             // :583-592
                        if (!p.sql || mse < thr || lastmse < thr) soft_count += soft_pending;
                        soft_pending = 0;
                    }
                }
            }
        }
        sig2_last = sig2;                                                 // :596
        if (!PRE) {   // this sample's mixed value enters the FIR ring (:453-456); finish the next output
            const double cre = c2_re * dval, cim = c2_im * dval;
            s_re[fir_pos * OQ_THREADS + lane] = cre; s_re[(fir_pos + OQ_NT1) * OQ_THREADS + lane] = cre;
            s_im[fir_pos * OQ_THREADS + lane] = cim; s_im[(fir_pos + OQ_NT1) * OQ_THREADS + lane] = cim;
            nfre += p.taps[54] * cre; nfim += p.taps[54] * cim;
            fir_pos++; if (fir_pos >= OQ_NT1) fir_pos = 0;
        }
        osc_next_frame(m2); osc_next_frame(st); osc_next_frame(sr);       // :600-603 (mixer_center advanced above)
        {
            const int t = osc_index(m2.ptr);
            if (t == m2_spec) { c2_re = n2_re; c2_im = n2_im; } else { c2_re = cos_t[t]; c2_im = sin_t[t]; }
        }
        { const int t = osc_index(st.ptr); cs_re = cos_t[t]; cs_im = sin_t[t]; }
        fre = nfre; fim = nfim;
        // ---- ring tile bookkeeping (warp-uniform)
        S++;
        if ((S & (OQ_T - 1)) == 0) {
            ring_store(rt);                                   // the finished tile goes back to HBM
            ring_dirty = false;
            rt++;
            if (S < S_end) {
                OQ_WAIT((int)(rt & 1));                       // next tile (requested a tile ago)
                ring_next_issued = false;
                if ((rt + 1) * OQ_T < S_end) {
                    bulk_wait_read_all();                     // the buffer being refilled must have been read out by its store
                    ring_load(rt + 1); ring_next_issued = true;
                }
            }
        }
    }
    // ---------------- drain the staging pipeline
    if (ring_dirty) ring_store(rt);
    if (ring_next_issued) OQ_WAIT((int)((rt + 1) & 1));
    if (pcm_next_issued) { OQ_WAIT(2 + ((pt + 1) & 1)); if (PRE) OQ_WAIT(4 + ((pt + 1) & 1)); }
    bulk_wait_all();

    // ---------------- store state
    LD(D_M2_PTR) = m2.ptr; LD(D_M2_STEP) = m2.step; LD(D_M2_FREQ) = m2.freq; LD(D_M2_LAST) = m2.last;
    LD(D_MC_PTR) = mc.ptr; LD(D_MC_STEP) = mc.step; LD(D_MC_FREQ) = mc.freq; LD(D_MC_LAST) = mc.last;
    LD(D_ST_PTR) = st.ptr; LD(D_ST_STEP) = st.step; LD(D_ST_FREQ) = st.freq; LD(D_ST_LAST) = st.last;
    LD(D_SR_PTR) = sr.ptr; LD(D_SR_STEP) = sr.step; LD(D_SR_FREQ) = sr.freq; LD(D_SR_LAST) = sr.last;
    LD(D_AGC_SUM) = agc_sum; LD(D_AGC_VAL) = agc_val;
    LD(D_EB_SUM1) = eb_sum1; LD(D_EB_SUM2) = eb_sum2; LD(D_EB_EBNO) = eb_ebno;
    LD(D_DLY_S0) = dly_s0;
    LD(D_DLY41_0) = d41_0; LD(D_DLY41_1) = d41_1; LD(D_DLY41_2) = d41_2;
    LD(D_DLY42_0) = d42_0; LD(D_DLY42_1) = d42_1; LD(D_DLY42_2) = d42_2;
    LD(D_DLY8_0) = d8_0; LD(D_DLY8_1) = d8_1; LD(D_DLY8_2) = d8_2;
    LD(D_RES_X1) = res.x1; LD(D_RES_X2) = res.x2; LD(D_RES_Y1) = res.y1; LD(D_RES_Y2) = res.y2;
    LD(D_LF_X1) = lf.x1; LD(D_LF_X2) = lf.x2; LD(D_LF_Y1) = lf.y1; LD(D_LF_Y2) = lf.y2;
    LD(D_SIG2L_RE) = sig2_last.x; LD(D_SIG2L_IM) = sig2_last.y;
    LD(D_PTD_RE) = pt_d.x; LD(D_PTD_IM) = pt_d.y;
    LD(D_MARG_SUM) = marg_sum; LD(D_MARG_VAL) = marg_val;
    LD(D_MSE_PM_SUM) = pm_sum; LD(D_MSE_MA_SUM) = ma_sum; LD(D_MSE) = mse;
    LD(D_LASTMSE) = lastmse;
    LD(D_SCAT0_RE) = sc0.x; LD(D_SCAT0_IM) = sc0.y; LD(D_SCAT1_RE) = sc1.x; LD(D_SCAT1_IM) = sc1.y;
    if (PRE) m2_freq_sum[ch] = m2sum;
    LI(I_YUI) = yui; LI(I_COUNTDOWN) = countdown; LI(I_COUNTDOWN2) = countdown2; LI(I_SIG2L_INIT) = sig2l_init;
    LI(I_MARG_POS) = marg_pos; LI(I_DT_POS) = dt_pos; LI(I_MSE_POS) = mse_pos;
    LI(I_SOFT_COUNT) = soft_count; LI(I_SOFT_PENDING) = soft_pending; LI(I_SOFT_OVERFLOW) = soft_overflow;
    LI(I_SIG_TRUE) = sig_true; LI(I_SIG_FALSE) = sig_false;
    for (int k = 0; k < OQ_NT1; k++) {
        p.fir_re[(size_t)k * cpad + ch] = s_re[k * OQ_THREADS + lane];
        p.fir_im[(size_t)k * cpad + ch] = s_im[k * OQ_THREADS + lane];
    }
}

int oqpsk_segment_launch(const DemodParams &p, const SegmentArgs &a, const int16_t *d_pcm, size_t stride, cudaStream_t s)
{
    const int grid = (p.n_channels + OQ_THREADS - 1) / OQ_THREADS;
    if (p.xpre) {
        const size_t smem = (size_t)OQ_SM_TOTAL_PRE;
        JB_CUDA(cudaFuncSetAttribute(oqpsk_segment_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        oqpsk_segment_kernel<true><<<grid, OQ_THREADS, smem, s>>>(p, a, d_pcm, stride, p.xpre, p.xstride, p.m2_freq_sum);
    } else {
        const size_t smem = (size_t)OQ_SM_TOTAL;
        JB_CUDA(cudaFuncSetAttribute(oqpsk_segment_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        oqpsk_segment_kernel<false><<<grid, OQ_THREADS, smem, s>>>(p, a, d_pcm, stride, nullptr, 0, nullptr);
    }
    JB_CUDA(cudaGetLastError());
    return 0;
}

#undef LD
#undef LI
} // namespace jb
