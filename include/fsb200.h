/*
 * fsb200.h -- C ABI of libfsb200.so : the B200 (sm_100a) implementation of the
 * Fractalshades per-pixel iteration hot path.
 *
 * Plain pointers and sizes only; no torch / numpy types.  Every entry point
 * names the reference interface it replaces (GBillotey/Fractalshades v1.2.1,
 * paths relative to src/fractalshades/):
 *
 *   fsb_std_run        <- Fractal.numba_cycle_call -> numba_cycles
 *                         (core.py:2022-2027, 2935-2963) for the standard
 *                         Mandelbrot / Burning-ship `calc_std_div`
 *   fsb_frame_create   <- PerturbationFractal.get_cycle_indep_args
 *                         (perturbation.py:431-562): per-frame tables; the
 *                         dZndc/dZndz paths (numba_dZndc_path[_BS],
 *                         numba_dZndz_path, perturbation.py:2282-2516) and the
 *                         BLA tree (numba_make_BLA[_BS], :1819-1973) are
 *                         computed by the library when not supplied
 *   fsb_frame_run      <- PerturbationFractal.numba_cycle_call ->
 *                         numba_cycles_perturb[_BS] (perturbation.py:414-428,
 *                         988-1045, 1406-1451)
 *   fsb_frame_run_device  same, with device-resident input/output (bench)
 *
 * Return codes: 0 ok, 1 = USER_INTERRUPTED (core.py:1280), < 0 error; the
 * message is available from fsb_last_error() (thread-local).
 * There is no CPU fallback: without a CUDA device every compute call fails.
 *
 * Array conventions are those of the reference (core.py:2044-2075): for npts
 * points, c_pix is complex128[npts]; Z is (n_Z, npts) row-major, complex128
 * for holomorphic models and float64 for the burning-ship family; U is
 * (n_U, npts) int32; stop_reason (1, npts) int8; stop_iter (1, npts) int32.
 * Xrange arrays (numpy_utils/xrange.py) are passed as a mantissa array plus an
 * int32 exponent array.
 */
#ifndef FSB200_H
#define FSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSB_OK 0
#define FSB_USER_INTERRUPTED 1

enum { FSB_MODEL_M2 = 0, FSB_MODEL_BS = 1 };

typedef struct fsb_stats {
    double kernel_ms;       /* CUDA-event time of the pixel kernel(s)            */
    double h2d_ms;          /* host->device copies (0 for *_device entry points) */
    double d2h_ms;          /* device->host copies                               */
    int64_t n_iter_exec;    /* executed full iterations (not BLA-skipped)        */
    int64_t n_bla_steps;    /* applied BLA steps                                 */
    int64_t n_rebase;       /* rebases (both kinds)                              */
    int64_t sum_stop_iter;  /* sum over pixels of stop_iter (effective iters)    */
    int64_t n_launches;     /* kernels launched by this call                     */
    int64_t n_iter_fast;    /* Xrange frames: iterations done on the fp64 fast path */
} fsb_stats;

/* ---- library / device ---------------------------------------------------- */
int fsb_device_count(void);            /* < 0 on CUDA error                     */
int fsb_init(int device);              /* bind this process to one GPU          */
void fsb_shutdown(void);
const char *fsb_last_error(void);
const char *fsb_build_info(void);      /* "fsb200 sm_100a fmad=on|off ..."      */
int fsb_device_info(char *name, int name_cap, int *sm_count, int64_t *mem_bytes);

/* pinned host memory and raw device memory (bench / tile scheduler) */
void *fsb_host_alloc(int64_t bytes);
void fsb_host_free(void *p);
void *fsb_dev_alloc(int64_t bytes);
void fsb_dev_free(void *p);
int fsb_memcpy_h2d(void *dst_dev, const void *src_host, int64_t bytes);
int fsb_memcpy_d2h(void *dst_host, const void *src_dev, int64_t bytes);
int fsb_dev_memset(void *dst_dev, int value, int64_t bytes);
/* write `bytes` of a scratch buffer: flushes the L2 between timed launches */
int fsb_flush_l2(void);

/* ---- pixel projection --------------------------------------
 * The numba closures `proj_impl` and `proj_dzndc_modifier` of cycle_indep_args
 * (perturbation.py:537-549; `Fractal.proj_impl`, core.py:2037-2041) as plain
 * parameters.  All zero = Cartesian without modifier.
 *   FSB_PROJ_EXPMAP      projection.Expmap.make_f_impl (projection.py:363-373):
 *                        pix -> exp(hmoy + pix_to_ht * pix)
 *   FSB_DZNDC_MOD_EXPMAP Expmap.make_dzndc_modifier (:455-471): the dz/dc rows
 *                        are multiplied by exp(Re(pix_to_ht * pix) + mod_param),
 *                        mod_param = hshift (hmoy, or hmoy - exp_step_hmoy in a
 *                        step of a large exponential zoom, :316-331)
 *   FSB_DZNDC_MOD_SEAM   Cartesian(expmap_seam).make_dzndc_modifier (:205-219):
 *                        |pix + 1e-6| * mod_param, mod_param = expmap_seam
 * exp / sin / cos are evaluated with a fixed operation sequence (error < 1
 * ulp), identical in both builds and in the CPU oracle. */
enum { FSB_PROJ_CARTESIAN = 0, FSB_PROJ_EXPMAP = 1 };
enum { FSB_DZNDC_MOD_NONE = 0, FSB_DZNDC_MOD_EXPMAP = 1, FSB_DZNDC_MOD_SEAM = 2 };
typedef struct fsb_proj_desc {
    int32_t kind;             /* FSB_PROJ_*                                    */
    int32_t dzndc_modifier;   /* FSB_DZNDC_MOD_* (perturbation frames only)    */
    double hmoy;              /* Expmap: (hmin + hmax) / 2                     */
    double pix_to_ht[2];      /* Expmap.pix_to_ht (:306-313), complex          */
    double mod_param;
} fsb_proj_desc;
/* the projection alone / the modifier alone on a list of pixels (tests) */
int fsb_proj_apply(const fsb_proj_desc *p, int64_t npts, const double *c_pix,
                   double *out_pix, double *out_modifier);

/* ---- standard escape-time loop ------------------------------------------- */
typedef struct fsb_std_desc {
    int32_t model;            /* FSB_MODEL_M2 / FSB_MODEL_BS                  */
    int32_t flavor;           /* BS family flavour 1..5 (burning_ship.py:63)  */
    double center_re, center_im;
    double dx;
    double lin_mat[4];        /* core.py:1470-1484                            */
    int64_t max_iter;
    double M_divergence_sq;
    double epsilon_stationnary_sq;
    int32_t calc_d2zndc2;
    int32_t calc_orbit;
    int64_t backshift;
    fsb_proj_desc proj;       /* dzndc_modifier must be 0 (core.py:2035)      */
    /* FSB_MODEL_M2 only.  0: Mandelbrot (models/mandelbrot_M2.py:310-334);
     * N in [2, 32]: Mandelbrot_N.calc_std_div, z -> z^N + c (models/
     * mandelbrot_Mn.py:63-350), z^(N-1) as a product chain; calc_orbit must be 0 */
    int32_t nexp;
    int32_t _pad;
} fsb_std_desc;

/* number of rows of Z for this description */
int fsb_std_nz(const fsb_std_desc *d);

int fsb_std_run(const fsb_std_desc *d, int64_t npts, const double *c_pix,
                double *Z, int8_t *stop_reason, int32_t *stop_iter,
                const volatile uint8_t *interrupted, fsb_stats *stats);
int fsb_std_run_device(const fsb_std_desc *d, int64_t npts,
                       const double *d_c_pix, double *d_Z,
                       int8_t *d_stop_reason, int32_t *d_stop_iter,
                       fsb_stats *stats);

/* Tile-list variants.  The point list is the concatenation of n_tiles
 * row-major tiles (tile k: tile_h[k] rows of tile_w[k] points, row 0 first --
 * the layout of Fractal.chunk_pixel_pos, core.py:1767-1830, and of the memmap
 * slabs, core.py:2362-2472); npts = sum(tile_w[k] * tile_h[k]).  Same results
 * as the flat calls; a warp then works on an 8 x 4 pixel patch instead of a
 * 32 x 1 strip, which keeps more lanes busy.  This is what the GPU tile
 * scheduler (Fractal.compute_rawdata_dev, core.py:2515-2554) calls. */
int fsb_std_run_tiles(const fsb_std_desc *d, int32_t n_tiles, const int32_t *tile_w,
                      const int32_t *tile_h, const double *c_pix, double *Z,
                      int8_t *stop_reason, int32_t *stop_iter,
                      const volatile uint8_t *interrupted, fsb_stats *stats);
int fsb_std_run_tiles_device(const fsb_std_desc *d, int32_t n_tiles, const int32_t *tile_w,
                             const int32_t *tile_h, const double *d_c_pix, double *d_Z,
                             int8_t *d_stop_reason, int32_t *d_stop_iter, fsb_stats *stats);

/* ---- perturbation frames ------------------------------------------------- */
typedef struct fsb_frame fsb_frame;   /* opaque, owns the device tables */

typedef struct fsb_frame_desc {
    int32_t model;            /* FSB_MODEL_M2 (holomorphic) / FSB_MODEL_BS    */
    int32_t flavor;
    int64_t L;                /* len(Zn_path)                                 */
    const double *Zn_path;    /* complex128[L]                                */
    /* Xrange reference points (perturbation.py:317-375); n_xr may be 0 */
    int64_t n_xr;
    const int32_t *ref_index_xr;
    const double *ref_xr;     /* M2: complex128[n_xr]; BS: refx float64[n_xr] */
    const int32_t *ref_xr_e;
    const double *refy_xr;    /* BS only                                      */
    const int32_t *refy_xr_e;
    int64_t ref_div_iter;
    int64_t ref_order;        /* 1<<62 when the reference is not a cycle      */
    double drift[2];          /* M2: complex mantissa; BS: {driftx, drifty}   */
    int32_t drift_e[2];       /* M2: drift_e[0];       BS: {ex, ey}           */
    double lin_scale;  int32_t lin_scale_e;  int32_t _pad0;
    double lin_mat[4];
    double kc;         int32_t kc_e;         int32_t _pad1;  /* BLA bound     */
    double scale_deriv; int32_t scale_deriv_e; int32_t _pad2; /* = dx_xr      */
    /* options of calc_std_div */
    int32_t xr_detect;        /* xr_detect_activated (perturbation.py:152)    */
    int32_t bla_activated;
    int32_t calc_dzndc;       /* M2: calc_dzndc ; BS: calc_hessian            */
    int32_t calc_dzndz;       /* interior_detect (M2 only)                    */
    int32_t calc_orbit;
    int32_t _pad3;
    int64_t backshift;
    int64_t max_iter;
    double M_divergence_sq;
    double epsilon_stationnary_sq;
    double BLA_eps;
    /* Optional caller-supplied tables (staged parity: feed the oracle's).
     * NULL => computed by the library. */
    const double *dZndc;      /* M2: complex128[L]; BS: float64[4][L]         */
    const int32_t *dZndc_e;   /* xr_detect only; M2: [L]; BS: [4][L]          */
    const double *dZndz;      /* complex128[L+1]                              */
    const int32_t *dZndz_e;
    const double *M_bla;      /* M2: complex128[2*bla_len]; BS: f64[8*bla_len]*/
    const double *r_bla;      /* float64[bla_len]                             */
    int64_t bla_len;
    int32_t stages_bla;
    /* FSB_MODEL_M2 only.  0: Perturbation_mandelbrot (models/mandelbrot_M2.py:
     * 591-627); N in [2, 32]: Perturbation_mandelbrot_N with exponent N -- the
     * binomial forms of p_iter_zn / p_iter_dzndc / p_iter_dzndz and dfdz =
     * N z^(N-1) (models/mandelbrot_Mn.py:628-742); calc_orbit must be 0 */
    int32_t nexp;
    fsb_proj_desc proj;
} fsb_frame_desc;

int fsb_frame_create(const fsb_frame_desc *desc, fsb_frame **out);
int fsb_frame_destroy(fsb_frame *f);
int fsb_frame_nz(const fsb_frame *f);          /* rows of Z                    */
int64_t fsb_frame_bla_len(const fsb_frame *f);
int fsb_frame_stages_bla(const fsb_frame *f);
double fsb_frame_setup_ms(const fsb_frame *f, int what); /* 0 upload, 1 dZndc, 2 BLA */
/* read back the per-frame tables (tests) */
int fsb_frame_get_bla(const fsb_frame *f, double *M_bla, double *r_bla);
int fsb_frame_get_dzndc(const fsb_frame *f, double *dZndc, int32_t *dZndc_e);
int fsb_frame_get_dzndz(const fsb_frame *f, double *dZndz, int32_t *dZndz_e);

int fsb_frame_run(fsb_frame *f, int64_t npts, const double *c_pix, double *Z,
                  int32_t *U, int8_t *stop_reason, int32_t *stop_iter,
                  const volatile uint8_t *interrupted, fsb_stats *stats);
int fsb_frame_run_device(fsb_frame *f, int64_t npts, const double *d_c_pix,
                         double *d_Z, int32_t *d_U, int8_t *d_stop_reason,
                         int32_t *d_stop_iter, fsb_stats *stats);

/* tile-list variants, see fsb_std_run_tiles */
int fsb_frame_run_tiles(fsb_frame *f, int32_t n_tiles, const int32_t *tile_w,
                        const int32_t *tile_h, const double *c_pix, double *Z, int32_t *U,
                        int8_t *stop_reason, int32_t *stop_iter,
                        const volatile uint8_t *interrupted, fsb_stats *stats);
int fsb_frame_run_tiles_device(fsb_frame *f, int32_t n_tiles, const int32_t *tile_w,
                               const int32_t *tile_h, const double *d_c_pix, double *d_Z,
                               int32_t *d_U, int8_t *d_stop_reason, int32_t *d_stop_iter,
                               fsb_stats *stats);

/* ---- post-processing of the raw fields on the device --------
 * Continuous iteration number (postproc.py:352-406, 1001-1009), distance
 * estimate (:684-731) and normal of the potential (:572-628, kind "potential")
 * for the "infinity" potential of the divergent models, computed from the raw
 * fields while they are still in HBM.  Outputs are float32 arrays of npts
 * values (the reference's settings.postproc_dtype; float64 when out_f64), NULL
 * = not wanted.  Values are only meaningful where stop_reason == 1. */
typedef struct fsb_postproc_desc {
    int32_t holomorphic;      /* 1: Z rows complex128 ; 0: float64 rows (xn, yn, dxnda..dyndb) */
    int32_t row_zn;           /* row of zn (xn) in Z                            */
    int32_t row_dzndc;        /* row of dzndc (dxnda); < 0: no derivative rows  */
    int32_t has_skew;
    double potential_d, potential_a_d, potential_M;   /* models' potential_* attributes */
    double floor_iter;        /* Continuous_iter_pp(floor_iter=...)             */
    double px_snap;           /* DEM_pp(px_snap=...), < 0: none                 */
    double skew[4];           /* Fractal.skew, row-major                        */
    int32_t out_f64;
    /* derivative of the projection applied to the dz/dc rows before DEM and normal
     * (Postproc.get_dzndc -> apply_df / apply_dfBS, postproc.py:184-206, 973-999, with
     * Expmap.df / dfBS, projection.py:375-453); needs the pixel offsets (c_pix):
     * 0 none (Cartesian); 1 exp(i Im(k pix)) (stepped flow, rotates_df); 2 exp(k pix);
     * 3 exp(Re(k pix)); k = Expmap.pix_to_ht                                         */
    int32_t df_kind;
    double df_k[2];
} fsb_postproc_desc;

/* fused: pixel kernels + post-processing in one call; Z / U never leave the
 * device, only the requested fields (and stop_reason / stop_iter when not
 * NULL) are copied back: 4-16 B per point instead of 41-60.  n_tiles > 0: tile
 * list (npts ignored); n_tiles = 0: flat list of npts points. */
int fsb_frame_run_pp(fsb_frame *f, int32_t n_tiles, const int32_t *tile_w,
                     const int32_t *tile_h, int64_t npts, const double *c_pix,
                     const fsb_postproc_desc *pp, void *nu, void *dem, void *normal_x,
                     void *normal_y, int8_t *stop_reason, int32_t *stop_iter,
                     const volatile uint8_t *interrupted, fsb_stats *stats);
/* ---- grid calls (GPU tile scheduler) --------------------------------------
 * Same as fsb_std_run_tiles / fsb_frame_run_tiles / fsb_frame_run_pp, but the
 * pixel offsets are given by per-tile axes instead of a complex128[npts]
 * array: `axes` holds, for each tile k in order, tile_w[k] x values then
 * tile_h[k] y values, and the point (row r, column col) of tile k is
 * x_k[col] + i y_k[r] -- exactly Fractal.chunk_pixel_pos without jitter
 * (core.py:1767-1830: meshgrid of two linspace axes; y already negated and
 * divided by xy_ratio).  The library expands them on the device (k_expand_grid):
 * a 4K frame sends 0.7 MB over PCIe instead of 133 MB.  Replaces the c_pix
 * argument built by Fractal.get_cycling_dep_args (core.py:2044-2075) for the
 * tiles dispatched by compute_rawdata_dev (core.py:2515-2554). */
int fsb_std_run_grid(const fsb_std_desc *d, int32_t n_tiles, const int32_t *tile_w,
                     const int32_t *tile_h, const double *axes, double *Z,
                     int8_t *stop_reason, int32_t *stop_iter,
                     const volatile uint8_t *interrupted, fsb_stats *stats);
int fsb_frame_run_grid(fsb_frame *f, int32_t n_tiles, const int32_t *tile_w,
                       const int32_t *tile_h, const double *axes, double *Z, int32_t *U,
                       int8_t *stop_reason, int32_t *stop_iter,
                       const volatile uint8_t *interrupted, fsb_stats *stats);
int fsb_frame_run_grid_pp(fsb_frame *f, int32_t n_tiles, const int32_t *tile_w,
                          const int32_t *tile_h, const double *axes,
                          const fsb_postproc_desc *pp, void *nu, void *dem, void *normal_x,
                          void *normal_y, int8_t *stop_reason, int32_t *stop_iter,
                          const volatile uint8_t *interrupted, fsb_stats *stats);
/* stand-alone: raw fields already on the device / on the host (n_rows rows of Z);
 * the *_proj forms take the pixel offsets a non-zero df_kind needs */
int fsb_postproc_run_device(const fsb_postproc_desc *pp, int64_t npts, int32_t n_rows,
                            const double *d_Z, const int32_t *d_stop_iter, void *d_nu,
                            void *d_dem, void *d_normal_x, void *d_normal_y);
int fsb_postproc_run(const fsb_postproc_desc *pp, int64_t npts, int32_t n_rows,
                     const double *Z, const int32_t *stop_iter, void *nu, void *dem,
                     void *normal_x, void *normal_y);

int fsb_postproc_run_proj_device(const fsb_postproc_desc *pp, int64_t npts, int32_t n_rows,
                                 const double *d_Z, const int32_t *d_stop_iter,
                                 const double *d_c_pix, void *d_nu, void *d_dem,
                                 void *d_normal_x, void *d_normal_y);
int fsb_postproc_run_proj(const fsb_postproc_desc *pp, int64_t npts, int32_t n_rows,
                          const double *Z, const int32_t *stop_iter, const double *c_pix,
                          void *nu, void *dem, void *normal_x, void *normal_y);

/* ---- field lines and Blinn shading (second part) -------------
 * Fieldlines_pp (postproc.py:409-531; Fieldlines_pp_infinity[_BS] :1038-1160): the
 * orbit is continued n_iter + 2 steps past the exit point with the model's
 * zn_iterate / xnyn_iterate (mandelbrot_M2.py:13-15, mandelbrot_Mn.py:13-17,
 * burning_ship.py:82-122), clamped on the circle of radius 1e5, and the sines of its
 * arguments are blended with Catmull-Rom weights of the fractional iteration
 * number.  Blinn_lighting.partial_shade (colors/layers.py:865-903) on the normal
 * of DEM_normal_pp scaled by sin(max_slope) (Color_layer.apply_shade :493-507):
 * the Lambert and the specular coefficient of each light source -- the two
 * per-pixel factors of the shading; the colour arithmetic that multiplies them
 * (XYZ of the base layer, k_diffuse, k_specular, light colour) stays with the
 * caller's layers.  "infinity" potential. */
#define FSB_PP_MAX_FL 32
#define FSB_PP_MAX_LIGHTS 4
typedef struct fsb_postproc_ext {
    int32_t fl_n_iter;        /* Fieldlines_pp(n_iter); 0: no field-lines output      */
    int32_t fl_row_orbit;     /* row of zn_orbit (xn_orbit) in Z; < 0: start from zn,
                                 not backward (postproc.py:470-481)                   */
    int32_t fl_backshift;     /* calc_orbit back-shift (context "backshift")          */
    int32_t fl_model;         /* 2..32: z -> z^N + c ; -1..-5: burning-ship flavour   */
    double fl_k[FSB_PP_MAX_FL];     /* k_arr   (geomspace(1, endpoint_k, n) / sum)    */
    double fl_phi[FSB_PP_MAX_FL];   /* phi_arr (default_rng(0).random(n) swirl pi)    */
    /* get_std_cpt (core.py:2781-2791, perturbation.py:197-208):
     * c = center + scale * lin_mat . pix                                             */
    double c_center[2], c_scale, c_lin_mat[4];
    int32_t n_lights;
    int32_t proj_kind;        /* FSB_PROJ_*: proj_impl of get_std_cpt (Expmap: exp(hmoy + k pix)) */
    double normal_coeff;      /* sin(Normal_map_layer.max_slope)                      */
    /* per light: LSx, LSy, LSz, half_x, half_y, half_z, shininess, (k_specular != 0) */
    double light[FSB_PP_MAX_LIGHTS][8];
    double proj_hmoy, proj_k[2];
} fsb_postproc_ext;
/* outputs (NULL = not wanted): fieldlines[npts]; shade[(2 n_lights) x npts] = rows
 * lambert_0, specular_0, lambert_1, ... ; element type as fsb_postproc_desc.out_f64.
 * c_pix: the pixel offsets of the points (complex128[npts]), needed by field lines. */
int fsb_postproc_ext_run_device(const fsb_postproc_desc *pp, const fsb_postproc_ext *ext,
                                int64_t npts, int32_t n_rows, const double *d_Z,
                                const int32_t *d_stop_iter, const double *d_c_pix,
                                void *d_fieldlines, void *d_shade);
int fsb_postproc_ext_run(const fsb_postproc_desc *pp, const fsb_postproc_ext *ext,
                         int64_t npts, int32_t n_rows, const double *Z,
                         const int32_t *stop_iter, const double *c_pix, void *fieldlines,
                         void *shade);
/* fsb_frame_run_grid_pp with the two extra outputs */
int fsb_frame_run_grid_pp_ext(fsb_frame *f, int32_t n_tiles, const int32_t *tile_w,
                              const int32_t *tile_h, const double *axes,
                              const fsb_postproc_desc *pp, const fsb_postproc_ext *ext,
                              void *nu, void *dem, void *normal_x, void *normal_y,
                              void *fieldlines, void *shade, int8_t *stop_reason,
                              int32_t *stop_iter, const volatile uint8_t *interrupted,
                              fsb_stats *stats);

/* ---- Xrange device arithmetic, exposed for unit tests --------------------
 * (mirror of the reference's tests/test_numba_xr.py; runs on the GPU)
 * op: 0 add, 1 sub, 2 mul ; complex operands (re, im interleaved) */
int fsb_xr_binop_c(int op, int64_t n, const double *a, const int32_t *ae,
                   const double *b, const int32_t *be, double *out,
                   int32_t *oute);
int fsb_xr_to_standard_c(int64_t n, const double *a, const int32_t *ae,
                         double *out);
int fsb_hypot_test(int64_t n, const double *x, const double *y, double *out);
/* FP64 FMA-pipe microbenchmark: returns measured TFLOP/s (dependent DFMA
 * chains, all SMs), used as the compute roofline denominator */
double fsb_fp64_peak_tflops(int iters);

#ifdef __cplusplus
}
#endif
#endif /* FSB200_H */
