/*
 * fsb_emul.cu -- HOST emulation of the lane state machine of k_perturb_m2_v2.
 *
 * TEST INFRASTRUCTURE ONLY (built and loaded by tests/emul_lib.py, never by the
 * product).  The event-driven pixel kernel is written as `__host__ __device__`
 * per-lane functions (fractalshades_b200/csrc/fsb_lane.cuh: lane_step,
 * m2_hot_iter); this file drives the very same functions one pixel at a time on
 * the CPU, so that the `-m "not gpu"` suite can check the state machine against
 * the oracle bit for bit (compiled with FSB_STRICT = the -fmad=false semantics)
 * and to tolerance (default build formulas with libm's exact fma).  What it
 * does not cover is the warp scheduling around the lanes (votes, refill): that
 * is exercised by the GPU tests.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../fractalshades_b200/csrc/fsb_lane.cuh"

using namespace fsb;

extern "C" {

typedef struct emul_frame {
    int64_t L;                 /* orbit length; Zn_path holds L elements      */
    const double *Zn_path;     /* complex128[L]                               */
    const double *dZndc;       /* complex128[L] or NULL                       */
    const int32_t *dZndc_e;    /* Xrange frames                               */
    int64_t n_xr;
    const int32_t *ref_index_xr;
    const double *ref_xr;      /* complex128[n_xr]                            */
    const int32_t *ref_xr_e;
    int64_t ref_div_iter, ref_order;
    double drift[2];
    int32_t drift_e, lin_scale_e;
    double lin_scale;
    double lin_mat[4];
    const double *M_bla, *r_bla;
    int64_t bla_len;
    int32_t stages_bla, xr_detect, bla_activated, calc_dzndc;
    int64_t max_iter;
    double M_divergence_sq;
} emul_frame;

int fsb_emul_strict(void)
{
#ifdef FSB_STRICT
    return 1;
#else
    return 0;
#endif
}

} /* extern "C" */

/* counters: n_exec, n_bla, n_reb, n_sum, n_fast, hot iterations, event visits, failed guards */
template <bool XR, bool DZNDC, bool BLA>
static void run(const FrameDev &f, int64_t npts, const C *c_pix, double *Z, int32_t *U,
                int8_t *sr, int32_t *si, unsigned long long *cnt)
{
    const double *T2 = f.T2;
    for (int64_t ipt = 0; ipt < npts; ipt++) {
        LaneM2 s;
        LaneCold k;
        memset(&s, 0, sizeof s);
        memset(&k, 0, sizeof k);
        lane_park(s, 0u);
        k.ipt = (int)ipt;
        s.flags = LF_INIT | LF_EV;
        for (;;) {
            lane_step<XR, DZNDC, BLA>(f, s, c_pix, Z, U, (signed char *)sr, si, cnt, 1, k);
            cnt[6]++;
            if (s.flags & LF_NEED) break;
        hot:
            bool ev, bad;
#if FSB_ZZ2
            double Cr, Ci;
            m2_hot_enter(s, Cr, Ci);
            for (;;) {
                const double *rec = T2 + 4 * (long long)s.w;
                m2_hot_iter_c<XR, DZNDC, BLA>(s, Cr, Ci, 2. * s.Zr, 2. * s.Zi, rec[2], rec[3]);
                m2_hot_flags_c<XR, DZNDC, BLA>(s, rec[0], rec[1], f.h3, f.esc_hi, ev, bad, Cr, Ci);
                s.Zr = 0.5 * rec[0]; s.Zi = 0.5 * rec[1];
                cnt[5]++;
                if (ev | bad) break;
            }
#else
            for (;;) {
                const double *rec = T2 + 4 * (long long)s.w;
                m2_hot_iter<XR, DZNDC, BLA>(s, s.Zr, s.Zi, rec[2], rec[3]);
                s.Zr = rec[0]; s.Zi = rec[1];
                m2_hot_flags<XR, DZNDC, BLA>(s, s.Zr, s.Zi, f.h3, f.esc_hi, ev, bad);
                cnt[5]++;
                if (ev | bad) break;
            }
#endif
#if FSB_SHORT_TRIP
            if (BLA && ev && !bad && m2_trip_is_short<XR>(f, s, s.Zr, s.Zi)) {
                if (lane_bla_short<XR, DZNDC>(f, s, k)) { s.flags &= ~LF_EV; goto hot; }
                s.flags |= LF_EV;
                continue;
            }
#endif
            if (XR && bad) { s.flags |= LF_EV | LF_BAD; cnt[7]++; }
            else s.flags |= LF_EV | LF_ITER;
        }
    }
}

extern "C" int fsb_emul_perturb_m2(const emul_frame *e, int64_t npts, const double *c_pix, double *Z,
                        int32_t *U, int8_t *sr, int32_t *si, uint64_t *counters)
{
    const int64_t L = e->L;
    if (L >= (1LL << 30) || e->max_iter >= (1LL << 30)) return -3;
    if (e->ref_order < (1LL << 30)) return -4;         /* periodic reference: not this kernel */
    FrameDev f;
    memset(&f, 0, sizeof f);
    /* orbit and derivative path with the zero pad element of the device tables */
    std::vector<C> Zn((size_t)L + 1, mkC(0., 0.)), dz, dz_std;
    std::vector<int> dze;
    memcpy(Zn.data(), e->Zn_path, (size_t)L * sizeof(C));
    f.L = L; f.Li = (int)L;
    f.Zn = Zn.data();
    const bool xr = e->xr_detect != 0, dc = e->calc_dzndc != 0;
    if (dc) {
        dz.assign((size_t)L + 1, mkC(0., 0.));
        memcpy(dz.data(), e->dZndc, (size_t)L * sizeof(C));
        f.dZndc = dz.data();
        if (xr) {
            dze.assign((size_t)L + 1, 0);
            memcpy(dze.data(), e->dZndc_e, (size_t)L * sizeof(int));
            f.dZndc_e = dze.data();
            dz_std.assign((size_t)L + 1, mkC(0., 0.));
            for (int64_t i = 0; i < L; i++)
                dz_std[(size_t)i] = mkC(flush_component(dz[(size_t)i].re, dze[(size_t)i]),
                                        flush_component(dz[(size_t)i].im, dze[(size_t)i]));
            f.dZndc_std = dz_std.data();
        }
    }
    f.n_xr = e->n_xr; f.n_xr_i = (int)e->n_xr;
    f.ref_index_xr = e->ref_index_xr;
    f.ref_xr = (const C *)e->ref_xr; f.ref_xr_e = e->ref_xr_e;
    f.ref_div_iter = e->ref_div_iter; f.ref_order = e->ref_order;
    const long long big = (1LL << 30);
    f.ref_div_i = (int)(e->ref_div_iter < big ? e->ref_div_iter : big);
    f.ref_div_m1_i = f.ref_div_i - 1;
    f.order_i = 0;
    long long fi = L;
    if (e->ref_div_iter < fi) fi = e->ref_div_iter;
    f.first_invalid_i = (int)fi;
    f.max_iter = e->max_iter; f.max_iter_i = (int)e->max_iter;
    f.drift[0] = e->drift[0]; f.drift[1] = e->drift[1];
    f.drift_e[0] = f.drift_e[1] = e->drift_e;
    f.lin_scale = e->lin_scale; f.lin_scale_e = e->lin_scale_e;
    for (int i = 0; i < 4; i++) f.lin_mat[i] = e->lin_mat[i];
    f.M_bla = e->M_bla; f.r_bla = e->r_bla;
    f.bla_len = e->bla_len; f.stages_bla = e->stages_bla;
    f.Mdiv_sq = e->M_divergence_sq;
    f.zstride = npts;
    f.esc_hi = esc_word(esc_hi_of(f.Mdiv_sq));
    const bool bla = e->bla_activated != 0 && e->stages_bla > 3 && e->bla_len > 0;
    std::vector<int> r2hi;
    if (bla) {                      /* as k_bla_r2hi */
        r2hi.resize((size_t)(3 * e->bla_len));
        for (int64_t i = 0; i < e->bla_len; i++) {
            r2hi[(size_t)i] = bla_r2hi(e->r_bla[i], 1.);
            r2hi[(size_t)(e->bla_len + i)] = bla_r2hi(e->r_bla[i], 0x1p600);
            r2hi[(size_t)(2 * e->bla_len + i)] = bla_rhi(e->r_bla[i]);
        }
        f.r2hi = r2hi.data();
        f.r2hi_up = r2hi.data() + e->bla_len;
        f.rhi = r2hi.data() + 2 * e->bla_len;
    }
    /* interleaved orbit table and pre-test words, as k_build_t2 / k_build_h3 */
    const int64_t n_rec = L + 16, n_h3 = FSB_H3_DIRECT ? n_rec : L / 8 + 4;
    std::vector<double> T2((size_t)n_rec * 4, 0.);
    std::vector<unsigned> h3((size_t)n_h3, 0u);
    const C *dsrc = dc ? (xr ? f.dZndc_std : f.dZndc) : nullptr;
    for (int64_t i = 0; i < n_rec; i++) {
        double *r = T2.data() + 4 * i;
        if (i + 1 < L + 1) { r[0] = mul_rn(FSB_ZSCALE, Zn[(size_t)i + 1].re); r[1] = mul_rn(FSB_ZSCALE, Zn[(size_t)i + 1].im); }
        if (dsrc && i < L + 1) { r[2] = mul_rn(FSB_TSCALE, dsrc[i].re); r[3] = mul_rn(FSB_TSCALE, dsrc[i].im); }
    }
    if (bla)
        for (int64_t j = 0; h3_slot((int)(8 * j)) < n_h3 && j < e->bla_len / 2; j++)
            if ((int64_t)f.first_invalid_i - 8 * j > 8) h3[(size_t)h3_slot((int)(8 * j))] = h3_word((unsigned)hi32(e->r_bla[2 * j]));
    f.T2 = T2.data();
    f.h3 = h3.data();

    unsigned long long cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const C *cp = (const C *)c_pix;
#define RUN(X, D, B) run<X, D, B>(f, npts, cp, Z, U, sr, si, cnt)
    if (xr) {
        if (dc) { if (bla) RUN(true, true, true); else RUN(true, true, false); }
        else { if (bla) RUN(true, false, true); else RUN(true, false, false); }
    } else {
        if (dc) { if (bla) RUN(false, true, true); else RUN(false, true, false); }
        else { if (bla) RUN(false, false, true); else RUN(false, false, false); }
    }
#undef RUN
    if (counters) for (int k = 0; k < 8; k++) counters[k] = cnt[k];
#ifdef FSB_DEBUG_WALK
    printf("BLA lookups %lld, passed stage 3 %lld, stage tests in the walk %lld\n", g_walk_lookups, g_walk_pass3, g_walk_tests);
#endif
    return 0;
}
