// K1b — fused continuous MSK demodulator segment kernel (600 / 1200 bps), one thread per channel.
//
// Replaces MskDemodulator::writeData (JAERO/mskdemodulator.cpp:313-488) and
// MskDemodulator::FreqOffsetEstimateSlot (:490-519): int16 -> coarse-estimator ring -> NCO mix ->
// half-sine matched filter 2*SPS taps x2 (:164-170, DSP.cpp:292-304) -> EbNo (DSP.cpp:493-505) -> AGC
// -> clip -> one-symbol delay (:384) -> resonator on |pt_msk| + T/2 quadrature + PLL with tanh
// weighting (:387-405) -> strobe (:408): carrier loop (:411-426), bias rotate (:429-431), MSE (:446-448),
// differential soft decode (:451-469, DSP.cpp:531-563), emit every 12 soft bits (:472-476).
// Same layout rules as K1a (oqpsk_demod.cu): lane = channel, lock-step ring positions.
#include "demod_device.cuh"

namespace jb {

#undef LD
#undef LI
static const int MSK_THREADS = 32;

#define LD(idx) p.D[(size_t)(idx) * p.cpad + ch]
#define LI(idx) p.I[(size_t)(idx) * p.cpad + ch]

// DiffDecode::UpdateSoft (DSP.cpp:531-563)
__device__ __forceinline__ double diff_update_soft(double &last, double soft)
{
    double r;
    if (soft < 0 && last < 0) { r = last; last = soft; }
    else if (soft > 0 && last > 0) { r = -last; last = soft; }
    else { r = fabs(last); last = soft; }
    return r;
}

__global__ void __launch_bounds__(MSK_THREADS)
msk_segment_kernel(const __grid_constant__ DemodParams p, const SegmentArgs a, const int16_t *__restrict__ pcm, size_t stride,
                   int d8_k, double d8_w)
{
    extern __shared__ double smem_d[];
    const int nt1 = p.ntaps + 1;                       // FIR ring length (DSP.cpp:277)
    double *s_re = smem_d;                             // [nt1][32]
    double *s_im = smem_d + (size_t)nt1 * MSK_THREADS;
    const int lane = threadIdx.x;
    const int ch = blockIdx.x * MSK_THREADS + lane;
    if (ch >= p.n_channels) return;

    Osc m2 = {LD(D_M2_PTR), LD(D_M2_STEP), LD(D_M2_FREQ), LD(D_M2_LAST)};
    Osc mc = {LD(D_MC_PTR), LD(D_MC_STEP), LD(D_MC_FREQ), LD(D_MC_LAST)};
    Osc st = {LD(D_ST_PTR), LD(D_ST_STEP), LD(D_ST_FREQ), LD(D_ST_LAST)};
    double agc_sum = LD(D_AGC_SUM), agc_val = LD(D_AGC_VAL);
    double eb_sum1 = LD(D_EB_SUM1), eb_sum2 = LD(D_EB_SUM2), eb_ebno = LD(D_EB_EBNO);
    Biquad res = {LD(D_RES_X1), LD(D_RES_X2), LD(D_RES_Y1), LD(D_RES_Y2)};
    double marg_sum = LD(D_MARG_SUM), marg_val = LD(D_MARG_VAL);
    double ma_sum = LD(D_MSE_MA_SUM), mse = LD(D_MSE);
    double2 sc0 = make_double2(LD(D_SCAT0_RE), LD(D_SCAT0_IM)), sc1 = make_double2(LD(D_SCAT1_RE), LD(D_SCAT1_IM));
    double diff_last = LD(D_DIFF_LAST);
    int countdown = LI(I_COUNTDOWN), dcd = LI(I_DCD);
    int marg_pos = LI(I_MARG_POS), dt_pos = LI(I_DT_POS), mse_pos = LI(I_MSE_POS);
    int soft_count = LI(I_SOFT_COUNT), soft_pending = LI(I_SOFT_PENDING), soft_overflow = LI(I_SOFT_OVERFLOW);
    int sig_true = LI(I_SIG_TRUE), sig_false = LI(I_SIG_FALSE);
    for (int k = 0; k < nt1; k++) {
        s_re[k * MSK_THREADS + lane] = p.fir_re[(size_t)k * p.cpad + ch];
        s_im[k * MSK_THREADS + lane] = p.fir_im[(size_t)k * p.cpad + ch];
    }

    // ---- FreqOffsetEstimateSlot (mskdemodulator.cpp:490-519)
    if (a.apply_cfe) {
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
                for (int j = 0; j < p.bb_len; j++) rowz[j] = make_double2(0.0, 0.0);    // :507
            }
        } else countdown = 4;
        if (mse > p.signalthreshold) { sig_false++; if (p.wire_sigstat) { { const int ln_ = LI(I_LOST_N); if (ln_ < LOST_CAP) p.lost_pos[(size_t)ln_ * p.cpad + ch] = LI(I_SOFT_COUNT); LI(I_LOST_N) = ln_ + 1; LI(I_DCD) = 0; } dcd = 0; } } else sig_true++;                       // :516-517
    }

    const int agc_len = p.agc_len, eb_len = p.ebno_len;
    const int ds_len = p.sps + 1, d8_len = d8_k + 1;
    int agc_pos = (int)(a.sample0 % agc_len), eb_pos = (int)(a.sample0 % eb_len);
    int fir_pos = (int)(a.sample0 % nt1), ds_pos = (int)(a.sample0 % ds_len), d8_pos = (int)(a.sample0 % d8_len);
    int bb_pos = a.bb_pos, coarse_counter = a.coarse_counter;
    const bool ebno_on = p.report_ebno != 0;
    const int16_t *row = pcm + (size_t)ch * stride;
    const int ntaps = p.ntaps;

    for (int i = a.i0; i < a.i1; i++) {
        const double dval = ((double)row[i]) / 32768.0;                                  // :322
        if (!(i == a.i0 && a.skip_a_first)) {                                            // :350-367
            if (coarse_counter >= p.Fs || !p.cpu_reduce) {
                const int t = osc_index(mc.ptr);
                p.bb[(size_t)ch * p.bb_len + bb_pos] = make_double2(p.cos_t[t] * dval, p.sin_t[t] * dval);
                bb_pos++; if (bb_pos >= p.bb_len) bb_pos = 0;
            }
        }
        if (i == a.i1 - 1 && a.stop_after_a) break;
        coarse_counter++;                                                                // :368

        const int t2 = osc_index(m2.ptr);
        const double cre = p.cos_t[t2] * dval, cim = p.sin_t[t2] * dval;                 // :369
        s_re[fir_pos * MSK_THREADS + lane] = cre; s_im[fir_pos * MSK_THREADS + lane] = cim;
        fir_pos++; if (fir_pos >= nt1) fir_pos = 0;
        double sre = 0, sim = 0;
        {
            int tp = fir_pos;
#pragma unroll 8
            for (int k = 0; k < ntaps; k++) {
                sre += p.taps[k] * s_re[tp * MSK_THREADS + lane];
                sim += p.taps[k] * s_im[tp * MSK_THREADS + lane];
                tp++; if (tp >= nt1) tp = 0;
            }
        }
        const double dabval = sqrt(sre * sre + sim * sim);                                // :372
        if (ebno_on) {                                                                    // MSKEbNoMeasure::Update (DSP.cpp:493-505)
            const size_t e = (size_t)eb_pos * p.cpad + ch;
            const double sq = dabval * dabval;
            eb_sum2 = eb_sum2 - p.ebno_e2[e]; eb_sum2 = eb_sum2 + fabs(sq); p.ebno_e2[e] = fabs(sq);
            eb_sum1 = eb_sum1 - p.ebno_e1[e]; eb_sum1 = eb_sum1 + fabs(dabval); p.ebno_e1[e] = fabs(dabval);
            const double e2val = eb_sum2 / ((double)eb_len), mean = eb_sum1 / ((double)eb_len);
            const double var = (e2val) - (mean * mean);
            const double alpha = sqrt(2.0) / mean;
            double tebno = 10.0 * (log10(2.0) - log10(((var * alpha * alpha) - 0.0085))) - 5.0;
            if (isnan(tebno)) tebno = 50;
            if (tebno > 50.0) tebno = 50;
            eb_ebno = eb_ebno * 0.8 + 0.2 * tebno;
        }
        eb_pos++; if (eb_pos >= eb_len) eb_pos = 0;
        {   // AGC::Update (DSP.cpp:370-379)
            const size_t e = (size_t)agc_pos * p.cpad + ch;
            agc_sum = agc_sum - p.agc_ring[e];
            agc_sum = agc_sum + fabs(dabval);
            p.agc_ring[e] = fabs(dabval);
            agc_pos++; if (agc_pos >= agc_len) agc_pos = 0;
            agc_val = 1.414213562 / fmax(agc_sum / ((double)agc_len), 0.000001);
            agc_val = fmax(agc_val, 0.000001);
        }
        double2 sig2 = make_double2(sre * agc_val, sim * agc_val);                        // :378
        const double abval = sqrt(sig2.x * sig2.x + sig2.y * sig2.y);                     // :381
        if (abval > 2.84) { const double g = (2.84 / abval); sig2 = make_double2(g * sig2.x, g * sig2.y); }

        // delayedsmpl.update_dont_touch(sig2): one symbol ago (:384, DSP.h:461-466)
        p.dsmpl_ring[(size_t)ds_pos * p.cpad + ch] = sig2;
        ds_pos++; if (ds_pos >= ds_len) ds_pos = 0;
        const double2 pt_d = p.dsmpl_ring[(size_t)ds_pos * p.cpad + ch];
        double2 pt_msk = make_double2(sig2.x, pt_d.y);                                    // :385

        double st_eta = biquad_update(res, hypot(pt_msk.x, pt_msk.y), p.res_a1, p.res_a2, p.res_b0, p.res_b1, p.res_b2);   // :387
        // delayt8.update(st_eta): Delay<double>(SPS/2) (DSP.h:357-374) as a ring with lock-step position
        double d8out;
        {
            p.dly8_ring[(size_t)d8_pos * p.cpad + ch] = st_eta;
            int io = d8_pos - d8_k; if (io < 0) io += d8_len;
            int in_ = io + 1; if (in_ >= d8_len) in_ = 0;
            const double older = p.dly8_ring[(size_t)io * p.cpad + ch];
            const double newer = p.dly8_ring[(size_t)in_ * p.cpad + ch];
            d8out = (d8_w * newer + (1.0 - d8_w) * older);
            d8_pos++; if (d8_pos >= d8_len) d8_pos = 0;
        }
        const int ts = osc_index(st.ptr);
        const double2 st_out = cmul(make_double2(p.cos_t[ts], p.sin_t[ts]), make_double2(st_eta, -d8out));   // :389-390
        const double st_angle_error = atan2_fast(st_out.y, st_out.x);                          // :392
        const double weighting = fabs(tanh(st_angle_error));                              // :395
        if (!dcd) osc_advance_fraction_of_wave(st, -(1.0 - weighting) * st_angle_error * (0.05 / 360.0));    // :397-405
        else osc_advance_fraction_of_wave(st, -(1.0 - weighting) * st_angle_error * (0.003 / 360.0));

        double frac;
        if (osc_have_passed_point(st, p.ee, frac)) {                                      // :408
            const double ct_xt = tanh(sig2.y) * sig2.x;
            const double ct_xt_d = tanh(pt_d.x) * pt_d.y;
            double ct_ec = ct_xt_d - ct_xt;
            if (ct_ec > M_PI) ct_ec = M_PI;
            if (ct_ec < -M_PI) ct_ec = -M_PI;
            if (ct_ec > M_PI_2) ct_ec = M_PI_2;
            if (ct_ec < -M_PI_2) ct_ec = -M_PI_2;
            double carrier_aggression = 12.0 * p.correctionfactor;                        // :422-426
            if (dcd) carrier_aggression = 8.0 * p.correctionfactor;
            osc_increase_phase_deg(m2, carrier_aggression * 1.0 * ct_ec);
            osc_set_freq(m2, (carrier_aggression * 0.01 * ct_ec) + m2.freq, p.Fs);
            {   // marg->UpdateSigned(ct_ec/2.0)  MA(SPS)  (:429)
                const size_t e = (size_t)marg_pos * p.cpad + ch;
                marg_sum = marg_sum - p.marg_ring[e];
                marg_sum = marg_sum + (ct_ec / 2.0);
                p.marg_ring[e] = (ct_ec / 2.0);
                marg_pos++; marg_pos %= p.marg_len;
                marg_val = marg_sum / ((double)p.marg_len);
            }
            {   // dt.update(pt_msk) (:430)
                p.dt_ring[(size_t)dt_pos * p.cpad + ch] = pt_msk;
                dt_pos++; dt_pos %= p.dt_len;
                pt_msk = p.dt_ring[(size_t)dt_pos * p.cpad + ch];
            }
            pt_msk = cmul(pt_msk, make_double2(cos(marg_val), sin(marg_val)));            // :431
            sc1 = sc0; sc0 = make_double2(pt_msk.x * 0.75, pt_msk.y * 0.75);                  // pointbuff (:440)
            {   // :446-448
                const double tda = (fabs((pt_msk).x * 0.75) - 1.0), tdb = (fabs((pt_msk).y * 0.75) - 1.0);
                const double v = (tda * tda) + (tdb * tdb);
                const size_t e = (size_t)mse_pos * p.cpad + ch;
                ma_sum = ma_sum - p.mse_ma[e]; ma_sum = ma_sum + fabs(v); p.mse_ma[e] = fabs(v);
                mse_pos++; mse_pos %= p.mse_len;
                mse = ma_sum / ((double)p.mse_len);
            }
            const double imagin = diff_update_soft(diff_last, pt_msk.y);                  // :451
            push_soft(p, ch, soft_count, soft_pending, soft_overflow, q_round((imagin) * 127.0 + 128.0));
            double real = diff_update_soft(diff_last, pt_msk.x);                          // :459
            real = -real;
            push_soft(p, ch, soft_count, soft_pending, soft_overflow, q_round((real) * 127.0 + 128.0));
            if (soft_pending >= 12) { soft_count += soft_pending; soft_pending = 0; }     // :472-476
        }
        osc_next_frame(m2); osc_next_frame(mc); osc_next_frame(st);                       // :480-483
    }

    LD(D_M2_PTR) = m2.ptr; LD(D_M2_STEP) = m2.step; LD(D_M2_FREQ) = m2.freq; LD(D_M2_LAST) = m2.last;
    LD(D_MC_PTR) = mc.ptr; LD(D_MC_STEP) = mc.step; LD(D_MC_FREQ) = mc.freq; LD(D_MC_LAST) = mc.last;
    LD(D_ST_PTR) = st.ptr; LD(D_ST_STEP) = st.step; LD(D_ST_FREQ) = st.freq; LD(D_ST_LAST) = st.last;
    LD(D_AGC_SUM) = agc_sum; LD(D_AGC_VAL) = agc_val;
    LD(D_EB_SUM1) = eb_sum1; LD(D_EB_SUM2) = eb_sum2; LD(D_EB_EBNO) = eb_ebno;
    LD(D_RES_X1) = res.x1; LD(D_RES_X2) = res.x2; LD(D_RES_Y1) = res.y1; LD(D_RES_Y2) = res.y2;
    LD(D_MARG_SUM) = marg_sum; LD(D_MARG_VAL) = marg_val;
    LD(D_MSE_MA_SUM) = ma_sum; LD(D_MSE) = mse; LD(D_DIFF_LAST) = diff_last;
    LD(D_SCAT0_RE) = sc0.x; LD(D_SCAT0_IM) = sc0.y; LD(D_SCAT1_RE) = sc1.x; LD(D_SCAT1_IM) = sc1.y;
    LI(I_COUNTDOWN) = countdown;
    LI(I_MARG_POS) = marg_pos; LI(I_DT_POS) = dt_pos; LI(I_MSE_POS) = mse_pos;
    LI(I_SOFT_COUNT) = soft_count; LI(I_SOFT_PENDING) = soft_pending; LI(I_SOFT_OVERFLOW) = soft_overflow;
    LI(I_SIG_TRUE) = sig_true; LI(I_SIG_FALSE) = sig_false;
    for (int k = 0; k < nt1; k++) {
        p.fir_re[(size_t)k * p.cpad + ch] = s_re[k * MSK_THREADS + lane];
        p.fir_im[(size_t)k * p.cpad + ch] = s_im[k * MSK_THREADS + lane];
    }
}

int msk_segment_launch(const DemodParams &p, const SegmentArgs &a, const int16_t *d_pcm, size_t stride, cudaStream_t s)
{
    const int grid = (p.n_channels + MSK_THREADS - 1) / MSK_THREADS;
    const size_t smem = (size_t)2 * (p.ntaps + 1) * MSK_THREADS * sizeof(double);
    JB_CUDA(cudaFuncSetAttribute(msk_segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    // Delay<double>(SPS/2) weight exactly as DSP.h:357-374 computes it at ring position 0
    const double fd = (p.sps) / 2.0;
    const int size = (int)ceil(fd) + 1;
    double dptr = 0.0 - fd;
    while (floor(dptr) < 0) dptr += (double)size;
    const double w = dptr - floor(dptr);
    msk_segment_kernel<<<grid, MSK_THREADS, smem, s>>>(p, a, d_pcm, stride, (int)ceil(fd), w);
    JB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace jb
