// K1b — warp-specialised continuous MSK demodulator segment kernel (600 / 1200 bps).
//
// Same arithmetic, statement for statement, as msk_segment_kernel (msk_demod.cu), i.e. as MskDemodulator::writeData
// (JAERO/mskdemodulator.cpp:313-488); same decomposition as the 10500 bps OQPSK pipeline (oqpsk_pipe.cu): the matched
// filter output of sample n excludes sample n (DSP.cpp:292-304), so FIR -> EbNo -> AGC -> clip -> one-symbol delay ->
// |pt_msk| -> resonator -> T/2 delay is feed-forward from the mixed samples up to n-1, and the symbol-rate tail after the
// carrier update feeds nothing back inside a call. Five warps per 32 channels (lane = channel in every warp):
//
//   warp F  2*SPS-tap half-sine matched filter of the mixed samples                                       -> sig2raw
//   warp E  EbNo + AGC running sums (TMA-staged ring tiles), AGC gain, clip, delayedsmpl, resonator, T/2   -> sig2, pt_d, st_eta, d8out
//   warp T  input (PCM tiles, coarse-estimator ring write), timing PLL (arg, tanh weighting, NCO nudge)    -> dval, strobe
//   warp K  carrier error, carrier NCO; mixes the next input sample into the FIR ring                      -> cval, (pt_msk, ct_ec)
//   warp S  marg MA(SPS), dt delay, bias rotate, MSE MA(600), differential soft decode, soft bits
#include "demod_device.cuh"

namespace jb {

static const int MP_THREADS = 160;
static const int MP_NBUF = 3;                       // ring-tile buffers (see oqpsk_pipe.cu: a tile is reloaded into the buffer stored a tile earlier)
static const int MP_HF = 16;                        // doubles per lane in a hand-off slot
// named barriers (0 is __syncthreads)
enum { MB_X = 1, MB_YT = 3, MB_Z = 5, MB_W = 7, MB_V = 9, MB_YK = 11, MB_U = 13 };

#undef LD
#undef LI
#define LD(idx) p.D[(size_t)(idx) * cpad + ch]
#define LI(idx) p.I[(size_t)(idx) * cpad + ch]

// DiffDecode::UpdateSoft (DSP.cpp:531-563)
__device__ __forceinline__ double mp_diff_update_soft(double &last, double soft)
{
    double r;
    if (soft < 0 && last < 0) { r = last; last = soft; }
    else if (soft > 0 && last > 0) { r = -last; last = soft; }
    else { r = fabs(last); last = soft; }
    return r;
}

// sum_{k<cnt} taps[k0+k] * ring[(start+k) % nt1] for both components, in tap order (DSP.cpp:296-303), as two
// contiguous runs of the ring
__device__ __forceinline__ void mp_fir_run(const DemodParams &p, const double *__restrict__ s_re, const double *__restrict__ s_im, int lane, int nt1, int start, int cnt,
                                           double &sre, double &sim)
{
    int k = 0, tp = start;
    const int first = min(cnt, nt1 - start);
#pragma unroll 8
    for (; k < first; k++, tp++) { sre += p.taps[k] * s_re[tp * 32 + lane]; sim += p.taps[k] * s_im[tp * 32 + lane]; }
    tp = 0;
#pragma unroll 8
    for (; k < cnt; k++, tp++) { sre += p.taps[k] * s_re[tp * 32 + lane]; sim += p.taps[k] * s_im[tp * 32 + lane]; }
}

__global__ void __launch_bounds__(MP_THREADS)
msk_pipe_kernel(const __grid_constant__ DemodParams p, const SegmentArgs a, const int16_t *__restrict__ pcm, size_t stride, int d8_k, double d8_w)
{
    extern __shared__ __align__(128) unsigned char mp_smem_raw[];
    const int ntaps = p.ntaps, nt1 = ntaps + 1;
    const int ds_len = p.sps + 1, d8_len = d8_k + 1;
    // shared memory map
    double *s_re = reinterpret_cast<double *>(mp_smem_raw);                        // [nt1][32]
    double *s_im = s_re + (size_t)nt1 * 32;
    double *t_agc = s_im + (size_t)nt1 * 32;                                       // [MP_NBUF][T][32]
    double *t_e1 = t_agc + MP_NBUF * OQ_T * 32;
    double *t_e2 = t_e1 + MP_NBUF * OQ_T * 32;
    double2 *s_ds = reinterpret_cast<double2 *>(t_e2 + MP_NBUF * OQ_T * 32);       // delayedsmpl [ds_len][32]
    double *s_d8 = reinterpret_cast<double *>(s_ds + (size_t)ds_len * 32);         // delayt8 [d8_len][32]
    double *hand = s_d8 + (size_t)d8_len * 32;                                     // [2][MP_HF][32]
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(hand + 2 * MP_HF * 32);   // ring tiles [MP_NBUF]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ch = blockIdx.x * 32 + lane;
    const bool live = ch < p.n_channels;
    const size_t cpad = p.cpad;
    if (threadIdx.x == 0) { for (int k = 0; k < MP_NBUF; k++) mbar_init(&bars[k], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();                                           // (0) mbarriers usable

    const int nB = (a.i1 - a.i0) - (a.stop_after_a ? 1 : 0);             // samples whose loop body runs in this launch
    const long long S0 = a.sample0;
    const double Fs = p.Fs;
    const double *__restrict__ cos_t = p.cos_t, *__restrict__ sin_t = p.sin_t;
    // hand-off slot: f 0,1 sig2raw (F->E) | 2..5 sig2, pt_d (E->K) | 6,7 st_eta, d8out (E->T) | 8,9 strobe, next input (T->K)
    //                | 10..13 flag, pt_msk.x, pt_msk.y, ct_ec (K->S) | 14 (slot 0) first input sample (T->K)
#define HAND(s, f) hand[((s) * MP_HF + (f)) * 32 + lane]
#define MP_WAIT(idx) do { mbar_wait(&bars[(idx)], (phases >> (idx)) & 1u); phases ^= (1u << (idx)); } while (0)

    // ======================================================================================= warp K: carrier loop
    if (warp == 3) {
        Osc m2 = {LD(D_M2_PTR), LD(D_M2_STEP), LD(D_M2_FREQ), LD(D_M2_LAST)};
        int dcd = LI(I_DCD);
        // ---- FreqOffsetEstimateSlot (mskdemodulator.cpp:490-519)
        if (a.apply_cfe) {
            Osc mc = {LD(D_MC_PTR), LD(D_MC_STEP), LD(D_MC_FREQ), LD(D_MC_LAST)};
            const double mse = LD(D_MSE);
            int countdown = LI(I_COUNTDOWN);
            const double est = p.cfe_est_out[ch];
            if ((mse > p.signalthreshold) && (fabs(m2.freq - (mc.freq + est)) > 0.0))      // :494-497
                osc_set_freq(m2, mc.freq + est, p.Fs);
            if ((p.afc) && (dcd) && (fabs(m2.freq - mc.freq) > 2.0)) {                      // :498-509
                if (countdown > 0) countdown--;
                else {
                    osc_set_freq(mc, m2.freq, p.Fs);
                    if (mc.freq < p.lockingbw / 2.0) osc_set_freq(mc, p.lockingbw / 2.0, p.Fs);
                    if (mc.freq > (p.Fs / 2.0 - p.lockingbw / 2.0)) osc_set_freq(mc, p.Fs / 2.0 - p.lockingbw / 2.0, p.Fs);
                    LI(I_EMPTYING) = 4; LI(I_ZERO_BB) = 1;                                   // bigchange()
                    double2 *rowz = p.bb + (size_t)ch * p.bb_len;
                    if (live) for (int j = 0; j < p.bb_len; j++) rowz[j] = make_double2(0.0, 0.0);    // :507
                    LD(D_MC_STEP) = mc.step; LD(D_MC_FREQ) = mc.freq;                        // warp T reloads mixer_center after the barrier
                }
            } else countdown = 4;
            if (mse > p.signalthreshold) { LI(I_SIG_FALSE) = LI(I_SIG_FALSE) + 1; if (p.wire_sigstat) { { const int ln_ = LI(I_LOST_N); if (ln_ < LOST_CAP) p.lost_pos[(size_t)ln_ * cpad + ch] = LI(I_SOFT_COUNT); LI(I_LOST_N) = ln_ + 1; LI(I_DCD) = 0; } dcd = 0; } }   // :516-517
            else LI(I_SIG_TRUE) = LI(I_SIG_TRUE) + 1;
            LI(I_COUNTDOWN) = countdown;
        }
        __syncthreads();                                       // (1) slot done, FIR ring resident
        if (nB > 0) {
            double c2_re, c2_im;
            { const int t = osc_index(m2.ptr); c2_re = cos_t[t]; c2_im = sin_t[t]; }
            int fir_pos = (int)(S0 % nt1);                     // slot of the sample being mixed
            {   // cval of the first sample (:369)
                const double dval = HAND(0, 14);
                s_re[fir_pos * 32 + lane] = c2_re * dval; s_im[fir_pos * 32 + lane] = c2_im * dval;
                fir_pos++; if (fir_pos >= nt1) fir_pos = 0;
                __threadfence_block();
                nb_arrive(MB_X + 0);                           // X_0
            }
            const double aggr = (dcd ? 8.0 : 12.0) * p.correctionfactor;      // :422-426
            for (int j = 0; j < nB; j++) {
                const int sl = j & 1;
                const int m2_spec = osc_next_index(m2);
                const double n2_re = cos_t[m2_spec], n2_im = sin_t[m2_spec];
                nb_sync(MB_YK + sl);                           // sig2, pt_d of this sample (warp E)
                const double2 sig2 = make_double2(HAND(sl, 2), HAND(sl, 3)), pt_d = make_double2(HAND(sl, 4), HAND(sl, 5));
                nb_sync(MB_U + sl);                            // strobe decision (warp T)
                const double strobe = HAND(sl, 8), dnext = HAND(sl, 9);
                double sy_flag = 0.0, sy_ec = 0.0;
                if (strobe != 0.0) {                                              // :408
                    const double ct_xt = tanh(sig2.y) * sig2.x;
                    const double ct_xt_d = tanh(pt_d.x) * pt_d.y;
                    double ct_ec = ct_xt_d - ct_xt;
                    if (ct_ec > M_PI) ct_ec = M_PI;
                    if (ct_ec < -M_PI) ct_ec = -M_PI;
                    if (ct_ec > M_PI_2) ct_ec = M_PI_2;
                    if (ct_ec < -M_PI_2) ct_ec = -M_PI_2;
                    osc_increase_phase_deg(m2, aggr * 1.0 * ct_ec);
                    osc_set_freq(m2, (aggr * 0.01 * ct_ec) + m2.freq, Fs);
                    sy_flag = 1.0; sy_ec = ct_ec;
                }
                osc_next_frame(m2);                                               // :480
                {
                    const int t = osc_index(m2.ptr);
                    if (t == m2_spec) { c2_re = n2_re; c2_im = n2_im; } else { c2_re = cos_t[t]; c2_im = sin_t[t]; }
                }
                if (j + 1 < nB) {   // the next sample's mixed value enters the FIR ring (:369-370)
                    s_re[fir_pos * 32 + lane] = c2_re * dnext; s_im[fir_pos * 32 + lane] = c2_im * dnext;
                    fir_pos++; if (fir_pos >= nt1) fir_pos = 0;
                    __threadfence_block();
                    nb_arrive(MB_X + ((j + 1) & 1));           // X_{j+1}
                }
                if (j >= 2) nb_sync(MB_V + sl);                // V_{j-2}: S has read slot sl
                HAND(sl, 10) = sy_flag; HAND(sl, 11) = sig2.x; HAND(sl, 12) = pt_d.y; HAND(sl, 13) = sy_ec;   // pt_msk=(sig2.re, pt_d.im) :385
                __threadfence_block();
                nb_arrive(MB_W + sl);                          // W_j
            }
            if (nB >= 2) nb_sync(MB_V + (nB & 1));             // V_{nB-2}
            nb_sync(MB_V + ((nB - 1) & 1));                    // V_{nB-1}
        }
        LD(D_M2_PTR) = m2.ptr; LD(D_M2_STEP) = m2.step; LD(D_M2_FREQ) = m2.freq; LD(D_M2_LAST) = m2.last;
    }
    // ======================================================================================= warp S: symbol-rate tail
    else if (warp == 4) {
        double marg_sum = LD(D_MARG_SUM), marg_val = LD(D_MARG_VAL);
        double ma_sum = LD(D_MSE_MA_SUM), mse = LD(D_MSE);
        double2 sc0 = make_double2(LD(D_SCAT0_RE), LD(D_SCAT0_IM)), sc1 = make_double2(LD(D_SCAT1_RE), LD(D_SCAT1_IM));
        double diff_last = LD(D_DIFF_LAST);
        int marg_pos = LI(I_MARG_POS), dt_pos = LI(I_DT_POS), mse_pos = LI(I_MSE_POS);
        int soft_count = LI(I_SOFT_COUNT), soft_pending = LI(I_SOFT_PENDING), soft_overflow = LI(I_SOFT_OVERFLOW);
        const int marg_len = p.marg_len, dt_len = p.dt_len, mse_len = p.mse_len;
        const double r_marg = 1.0 / ((double)marg_len), r_mse = 1.0 / ((double)mse_len);
        __syncthreads();                                       // (1)
        for (int j = 0; j < nB; j++) {
            const int sl = j & 1;
            nb_sync(MB_W + sl);                                // W_j
            const double fl = HAND(sl, 10), fx = HAND(sl, 11), fy = HAND(sl, 12), fec = HAND(sl, 13);
            __threadfence_block();
            nb_arrive(MB_V + sl);                              // V_j: slot read
            if (fl != 0.0) {
                double2 pt_msk = make_double2(fx, fy);
                const double ct_ec = fec;
                {   // marg->UpdateSigned(ct_ec/2.0)  MA(SPS)  (:429)
                    const size_t e = (size_t)marg_pos * cpad + ch;
                    marg_sum = marg_sum - p.marg_ring[e];
                    marg_sum = marg_sum + (ct_ec / 2.0);
                    p.marg_ring[e] = (ct_ec / 2.0);
                    marg_pos++; marg_pos %= marg_len;
                    marg_val = div_exact(marg_sum, (double)marg_len, r_marg);
                }
                {   // dt.update(pt_msk) (:430)
                    p.dt_ring[(size_t)dt_pos * cpad + ch] = pt_msk;
                    dt_pos++; dt_pos %= dt_len;
                    pt_msk = p.dt_ring[(size_t)dt_pos * cpad + ch];
                }
                pt_msk = cmul(pt_msk, make_double2(cos(marg_val), sin(marg_val)));            // :431
                sc1 = sc0; sc0 = make_double2(pt_msk.x * 0.75, pt_msk.y * 0.75);                  // pointbuff (:440)
                {   // :446-448
                    const double tda = (fabs((pt_msk).x * 0.75) - 1.0), tdb = (fabs((pt_msk).y * 0.75) - 1.0);
                    const double v = (tda * tda) + (tdb * tdb);
                    const size_t e = (size_t)mse_pos * cpad + ch;
                    ma_sum = ma_sum - p.mse_ma[e]; ma_sum = ma_sum + fabs(v); p.mse_ma[e] = fabs(v);
                    mse_pos++; mse_pos %= mse_len;
                    mse = div_exact(ma_sum, (double)mse_len, r_mse);
                }
                const double imagin = mp_diff_update_soft(diff_last, pt_msk.y);               // :451
                if (live) push_soft(p, ch, soft_count, soft_pending, soft_overflow, q_round((imagin) * 127.0 + 128.0));
                double real = mp_diff_update_soft(diff_last, pt_msk.x);                       // :459
                real = -real;
                if (live) push_soft(p, ch, soft_count, soft_pending, soft_overflow, q_round((real) * 127.0 + 128.0));
                if (soft_pending >= 12) { soft_count += soft_pending; soft_pending = 0; }     // :472-476
            }
        }
        LD(D_MARG_SUM) = marg_sum; LD(D_MARG_VAL) = marg_val;
        LD(D_MSE_MA_SUM) = ma_sum; LD(D_MSE) = mse; LD(D_DIFF_LAST) = diff_last;
        LD(D_SCAT0_RE) = sc0.x; LD(D_SCAT0_IM) = sc0.y; LD(D_SCAT1_RE) = sc1.x; LD(D_SCAT1_IM) = sc1.y;
        LI(I_MARG_POS) = marg_pos; LI(I_DT_POS) = dt_pos; LI(I_MSE_POS) = mse_pos;
        LI(I_SOFT_COUNT) = soft_count; LI(I_SOFT_PENDING) = soft_pending; LI(I_SOFT_OVERFLOW) = soft_overflow;
    }
    // ======================================================================================= warp T: input + symbol-timing PLL
    else if (warp == 2) {
        Osc st = {LD(D_ST_PTR), LD(D_ST_STEP), LD(D_ST_FREQ), LD(D_ST_LAST)};
        const int16_t *row = pcm + (size_t)ch * stride;
        // PCM: vector loads of the lane's own channel row, 8 samples at a time, one block ahead of use (see oqpsk_pipe.cu)
        const int4 *row4 = reinterpret_cast<const int4 *>(row);
        auto ld_blk = [&](int blk) -> int4 {
            return (live && (long long)blk * 8 < (long long)stride) ? __ldg(row4 + blk) : make_int4(0, 0, 0, 0);
        };
        int pk_blk = a.i0 >> 3;
        int4 pk = ld_blk(pk_blk), pk_next = ld_blk(pk_blk + 1);
        auto dval_at = [&](int ii) -> double {                    // ((double)*ptr)/32768.0 (:322); ii advances by one per call
            if ((ii >> 3) != pk_blk) { pk_blk = ii >> 3; pk = pk_next; pk_next = ld_blk(pk_blk + 1); }
            const int k = ii & 7;
            const int w = (k < 2) ? pk.x : (k < 4) ? pk.y : (k < 6) ? pk.z : pk.w;
            int v = (k & 1) ? (w >> 16) : (int)(short)(w & 0xffff);
            if (!live) v = 0;
            return ((double)v) / 32768.0;
        };
        double dcur = dval_at(a.i0);
        HAND(0, 14) = dcur;
        __syncthreads();                                       // (1) the slot may have re-centred mixer_center or (wired) cleared DCD
        const int dcd = LI(I_DCD);
        Osc mc = {LD(D_MC_PTR), LD(D_MC_STEP), LD(D_MC_FREQ), LD(D_MC_LAST)};
        int bb_pos = a.bb_pos, coarse_counter = a.coarse_counter;
        double2 *bb_row = p.bb + (size_t)ch * p.bb_len;
        const int bbn = p.bb_len;
        const bool cpu_reduce = p.cpu_reduce != 0;
        const double ee = p.ee;
        const double gain = dcd ? (0.003 / 360.0) : (0.05 / 360.0);           // :397-405
        double cs_re, cs_im, cc_re, cc_im;
        { const int t = osc_index(st.ptr); cs_re = cos_t[t]; cs_im = sin_t[t]; }
        { const int t = osc_index(mc.ptr); cc_re = cos_t[t]; cc_im = sin_t[t]; }
        for (int i = a.i0; i < a.i1; i++) {
            const int j = i - a.i0, sl = j & 1;
            if (!(i == a.i0 && a.skip_a_first)) {                                            // :350-367
                if (coarse_counter >= Fs || !cpu_reduce) {
                    if (live) bb_row[bb_pos] = make_double2(cc_re * dcur, cc_im * dcur);
                    bb_pos++; if (bb_pos >= bbn) bb_pos = 0;
                }
            }
            if (i == a.i1 - 1 && a.stop_after_a) break;
            coarse_counter++;                                                                // :368
            osc_next_frame(mc);                                                              // :481
            { const int t = osc_index(mc.ptr); cc_re = cos_t[t]; cc_im = sin_t[t]; }
            const double dnxt = (i + 1 < a.i1) ? dval_at(i + 1) : 0.0;
            nb_sync(MB_YT + sl);                               // st_eta, d8out of this sample (warp E)
            const double st_eta = HAND(sl, 6), d8out = HAND(sl, 7);
            const double2 st_out = cmul(make_double2(cs_re, cs_im), make_double2(st_eta, -d8out));   // :389-390
            const double st_angle_error = atan2_fast(st_out.y, st_out.x);                          // :392
            const double weighting = fabs(tanh(st_angle_error));                              // :395
            osc_advance_fraction_of_wave(st, -(1.0 - weighting) * st_angle_error * gain);
            double frac = 0.0;
            const bool strobe = osc_have_passed_point(st, ee, frac);                          // :408
            HAND(sl, 8) = strobe ? 1.0 : 0.0; HAND(sl, 9) = dnxt;
            __threadfence_block();
            nb_arrive(MB_U + sl);
            osc_next_frame(st);                                                               // :483
            { const int t = osc_index(st.ptr); cs_re = cos_t[t]; cs_im = sin_t[t]; }
            dcur = dnxt;
        }
        LD(D_ST_PTR) = st.ptr; LD(D_ST_STEP) = st.step; LD(D_ST_FREQ) = st.freq; LD(D_ST_LAST) = st.last;
        LD(D_MC_PTR) = mc.ptr; LD(D_MC_STEP) = mc.step; LD(D_MC_FREQ) = mc.freq; LD(D_MC_LAST) = mc.last;
    }
    // ======================================================================================= warp E: envelope chain
    else if (warp == 1) {
        double agc_sum = LD(D_AGC_SUM), agc_val = LD(D_AGC_VAL);
        double eb_sum1 = LD(D_EB_SUM1), eb_sum2 = LD(D_EB_SUM2), eb_ebno = LD(D_EB_EBNO);
        Biquad res = {LD(D_RES_X1), LD(D_RES_X2), LD(D_RES_Y1), LD(D_RES_Y2)};
        const int agc_len = p.agc_len, eb_len = p.ebno_len;
        const bool ebno_on = p.report_ebno != 0;
        const double r_agc = 1.0 / ((double)agc_len), r_eb = 1.0 / ((double)eb_len);
        const double res_a1 = p.res_a1, res_a2 = p.res_a2, res_b0 = p.res_b0, res_b1 = p.res_b1, res_b2 = p.res_b2;
        long long S = S0;
        int ds_pos = (int)(S % ds_len), d8_pos = (int)(S % d8_len);
        const long long S_end = S + nB;
        const int eb_from_j = (a.i1 - a.i0) - OQ_EBNO_TAIL;
        for (int k = 0; k < ds_len; k++) s_ds[k * 32 + lane] = p.dsmpl_ring[(size_t)k * cpad + ch];
        for (int k = 0; k < d8_len; k++) s_d8[k * 32 + lane] = p.dly8_ring[(size_t)k * cpad + ch];
        __syncthreads();                                       // (1)
        if (nB > 0) {
            // ring layout of this kernel: [cta][slot][32 lanes]; a tile (32 slots) is one contiguous 8 KB block moved by one bulk copy
            auto ring_tile = [&](double *ring, int len, long long tile) -> double * {
                return ring + ((size_t)blockIdx.x * len + (size_t)((tile * OQ_T) % len)) * 32;
            };
            const unsigned ring_tx = (ebno_on ? 3u : 1u) * OQ_SM_RING;
            auto ring_load = [&](long long tile) {
                const int b = (int)(tile % MP_NBUF);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    mbar_expect_tx(&bars[b], ring_tx);
                    bulk_g2s(t_agc + b * OQ_T * 32, ring_tile(p.agc_ring, agc_len, tile), OQ_SM_RING, &bars[b]);
                    if (ebno_on) {
                        bulk_g2s(t_e1 + b * OQ_T * 32, ring_tile(p.ebno_e1, eb_len, tile), OQ_SM_RING, &bars[b]);
                        bulk_g2s(t_e2 + b * OQ_T * 32, ring_tile(p.ebno_e2, eb_len, tile), OQ_SM_RING, &bars[b]);
                    }
                }
            };
            auto ring_store = [&](long long tile) {
                const int b = (int)(tile % MP_NBUF);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    bulk_s2g(ring_tile(p.agc_ring, agc_len, tile), t_agc + b * OQ_T * 32, OQ_SM_RING);
                    if (ebno_on) {
                        bulk_s2g(ring_tile(p.ebno_e1, eb_len, tile), t_e1 + b * OQ_T * 32, OQ_SM_RING);
                        bulk_s2g(ring_tile(p.ebno_e2, eb_len, tile), t_e2 + b * OQ_T * 32, OQ_SM_RING);
                    }
                    bulk_commit();
                }
            };
            unsigned phases = 0u;
            long long rt = S / OQ_T;                                  // current ring tile
            bool ring_next_issued = false, ring_dirty = false;
            // NOTE: the ring tiles assume agc_len and ebno_len are multiples of the tile length (48000 / 96000 / 24000 ...:
            // all multiples of 32 for the sample rates the reference uses; checked on the host)
            ring_load(rt);
            if ((rt + 1) * OQ_T < S_end) { ring_load(rt + 1); ring_next_issued = true; }
            MP_WAIT((int)(rt % MP_NBUF));
            for (int j = 0; j < nB; j++) {
                const int sl = j & 1;
                const int ro = (int)(S & (OQ_T - 1));
                const int rslot = (((int)(rt % MP_NBUF)) * OQ_T + ro) * 32 + lane;
                nb_sync(MB_Z + sl);                            // Z_j: matched filter output of this sample
                const double sre = HAND(sl, 0), sim = HAND(sl, 1);
                const double dabval = sqrt(sre * sre + sim * sim);                                // :372
                if (ebno_on) {                                                                    // MSKEbNoMeasure::Update (DSP.cpp:493-505)
                    const double sq = dabval * dabval;
                    eb_sum2 = eb_sum2 - t_e2[rslot]; eb_sum2 = eb_sum2 + fabs(sq); t_e2[rslot] = fabs(sq);
                    eb_sum1 = eb_sum1 - t_e1[rslot]; eb_sum1 = eb_sum1 + fabs(dabval); t_e1[rslot] = fabs(dabval);
                    // observable only: the smoothed read-out forgets its past by 0.8^k, evaluate it over the launch's tail
                    if (j >= eb_from_j) {
                        const double e2val = div_exact(eb_sum2, (double)eb_len, r_eb), mean = div_exact(eb_sum1, (double)eb_len, r_eb);
                        const double var = (e2val) - (mean * mean);
                        const double alpha = sqrt(2.0) / mean;
                        double tebno = 10.0 * (log10(2.0) - log10(((var * alpha * alpha) - 0.0085))) - 5.0;
                        if (isnan(tebno)) tebno = 50;
                        if (tebno > 50.0) tebno = 50;
                        eb_ebno = eb_ebno * 0.8 + 0.2 * tebno;
                    }
                }
                {   // AGC::Update (DSP.cpp:370-379)
                    agc_sum = agc_sum - t_agc[rslot];
                    agc_sum = agc_sum + fabs(dabval);
                    t_agc[rslot] = fabs(dabval);
                    ring_dirty = true;
                    agc_val = 1.414213562 / fmax(div_exact(agc_sum, (double)agc_len, r_agc), 0.000001);
                    agc_val = fmax(agc_val, 0.000001);
                }
                double2 sig2 = make_double2(sre * agc_val, sim * agc_val);                        // :378
                const double abval = sqrt(sig2.x * sig2.x + sig2.y * sig2.y);                     // :381
                if (abval > 2.84) { const double g = (2.84 / abval); sig2 = make_double2(g * sig2.x, g * sig2.y); }
                // delayedsmpl.update_dont_touch(sig2): one symbol ago (:384, DSP.h:461-466)
                s_ds[ds_pos * 32 + lane] = sig2;
                ds_pos++; if (ds_pos >= ds_len) ds_pos = 0;
                const double2 pt_d = s_ds[ds_pos * 32 + lane];
                const double2 pt_msk = make_double2(sig2.x, pt_d.y);                              // :385
                const double st_eta = biquad_update(res, hypot(pt_msk.x, pt_msk.y), res_a1, res_a2, res_b0, res_b1, res_b2);   // :387
                double d8out;                                                                     // delayt8.update(st_eta): Delay<double>(SPS/2)
                {
                    s_d8[d8_pos * 32 + lane] = st_eta;
                    int io = d8_pos - d8_k; if (io < 0) io += d8_len;
                    int in_ = io + 1; if (in_ >= d8_len) in_ = 0;
                    const double older = s_d8[io * 32 + lane], newer = s_d8[in_ * 32 + lane];
                    d8out = (d8_w * newer + (1.0 - d8_w) * older);
                    d8_pos++; if (d8_pos >= d8_len) d8_pos = 0;
                }
                HAND(sl, 2) = sig2.x; HAND(sl, 3) = sig2.y; HAND(sl, 4) = pt_d.x; HAND(sl, 5) = pt_d.y; HAND(sl, 6) = st_eta; HAND(sl, 7) = d8out;
                __threadfence_block();
                nb_arrive(MB_YT + sl);                         // timing inputs -> warp T
                nb_arrive(MB_YK + sl);                         // sig2, pt_d -> warp K
                S++;
                if ((S & (OQ_T - 1)) == 0) {
                    ring_store(rt);
                    ring_dirty = false;
                    rt++;
                    if (S < S_end) {
                        MP_WAIT((int)(rt % MP_NBUF));
                        ring_next_issued = false;
                        if ((rt + 1) * OQ_T < S_end) {
                            bulk_wait_read_1();
                            ring_load(rt + 1); ring_next_issued = true;
                        }
                    }
                }
            }
            if (ring_dirty) ring_store(rt);
            if (ring_next_issued) MP_WAIT((int)((rt + 1) % MP_NBUF));
            bulk_wait_all();
        }
        for (int k = 0; k < ds_len; k++) p.dsmpl_ring[(size_t)k * cpad + ch] = s_ds[k * 32 + lane];
        for (int k = 0; k < d8_len; k++) p.dly8_ring[(size_t)k * cpad + ch] = s_d8[k * 32 + lane];
        LD(D_AGC_SUM) = agc_sum; LD(D_AGC_VAL) = agc_val;
        LD(D_EB_SUM1) = eb_sum1; LD(D_EB_SUM2) = eb_sum2; LD(D_EB_EBNO) = eb_ebno;
        LD(D_RES_X1) = res.x1; LD(D_RES_X2) = res.x2; LD(D_RES_Y1) = res.y1; LD(D_RES_Y2) = res.y2;
    }
    // ======================================================================================= warp F: matched filter
    else {
        for (int k = 0; k < nt1; k++) {
            s_re[k * 32 + lane] = p.fir_re[(size_t)k * cpad + ch];
            s_im[k * 32 + lane] = p.fir_im[(size_t)k * cpad + ch];
        }
        __syncthreads();                                       // (1)
        // output j (:370) = sum over the ntaps mixed samples older than sample i0+j; the newest of them (slot `tail`) is produced
        // by warp K one sample earlier, the ntaps-1 older terms are summed ahead of that (same order as DSP.cpp:296-303)
        int tail = (int)((S0 + nt1 - 1) % nt1);
        double nfre = 0, nfim = 0;
        auto older = [&]() { nfre = 0; nfim = 0; int st0 = tail + 2; if (st0 >= nt1) st0 -= nt1; mp_fir_run(p, s_re, s_im, lane, nt1, st0, ntaps - 1, nfre, nfim); };
        if (nB > 0) older();
        for (int j = 0; j < nB; j++) {
            if (j > 0) nb_sync(MB_X + ((j - 1) & 1));         // X_{j-1}
            nfre += p.taps[ntaps - 1] * s_re[tail * 32 + lane]; nfim += p.taps[ntaps - 1] * s_im[tail * 32 + lane];
            const int sl = j & 1;
            HAND(sl, 0) = nfre; HAND(sl, 1) = nfim;
            __threadfence_block();
            nb_arrive(MB_Z + sl);                              // Z_j
            tail++; if (tail >= nt1) tail = 0;
            if (j + 1 < nB) older();
        }
        if (nB > 0) nb_sync(MB_X + ((nB - 1) & 1));           // X_{nB-1}: pair the last arrival of warp K
    }
    __syncthreads();                                           // (2) every warp is done with the FIR ring
    for (int k = warp; k < nt1; k += 5) {
        p.fir_re[(size_t)k * cpad + ch] = s_re[k * 32 + lane];
        p.fir_im[(size_t)k * cpad + ch] = s_im[k * 32 + lane];
    }
#undef MP_WAIT
#undef HAND
}

int msk_pipe_launch(const DemodParams &p, const SegmentArgs &a, const int16_t *d_pcm, size_t stride, cudaStream_t s)
{
    const int grid = (p.n_channels + 31) / 32;
    // Delay<double>(SPS/2) weight exactly as DSP.h:357-374 computes it at ring position 0
    const double fd = (p.sps) / 2.0;
    const int size = (int)ceil(fd) + 1;
    double dptr = 0.0 - fd;
    while (floor(dptr) < 0) dptr += (double)size;
    const double w = dptr - floor(dptr);
    const int d8_k = (int)ceil(fd);
    const size_t smem = (size_t)2 * (p.ntaps + 1) * 32 * 8 + 3 * MP_NBUF * OQ_SM_RING + (size_t)(p.sps + 1) * 32 * 16 + (size_t)(d8_k + 1) * 32 * 8 +
                        (size_t)2 * MP_HF * 32 * 8 + 64;
    JB_CUDA(cudaFuncSetAttribute(msk_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    msk_pipe_kernel<<<grid, MP_THREADS, smem, s>>>(p, a, d_pcm, stride, d8_k, w);
    JB_CUDA(cudaGetLastError());
    return 0;
}

#undef LD
#undef LI
} // namespace jb
