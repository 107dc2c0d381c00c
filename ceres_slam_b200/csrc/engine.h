// Host-side state of one bundle-adjustment problem on one B200, and the kernel launch surface.
#pragma once
#include <stdint.h>

#include <memory>
#include <string>
#include <vector>

#include "../../include/cslam_b200.h"
#include "closed_form.h"
#include "common.cuh"
#include "dogleg_host.h"

namespace cslam {

// Scalar slots accumulated on the device (doubles), one 32-slot block per use.
enum Scal {
    SC_COST = 0,         // 1/2 sum rho(|r|^2) at x (schur / colnorm pass)
    SC_MODEL = 1,        // model cost change  -(J s).(r + J s / 2)
    SC_CAND_COST = 2,    // cost at the candidate point
    SC_STEP_NORM2 = 3,   // |x - x_cand|^2 (ambient)
    SC_XNORM2 = 4,       // |x|^2 over free blocks (ambient) at the candidate
    SC_INVALID = 5,      // count of point blocks whose V was not positive definite
    SC_GRADMAX = 6,      // |x - Plus(x,-g)|_inf (bit-pattern max)
    SC_XNORM2_CUR = 7,   // |x|^2 at the current point (initial pass)
    SC_NONFINITE = 8,    // non-finite step entries
    SC_LS_GY = 9,        // g . y (scaled coordinates): the line search's directional derivative is -g.y
    SC_LS_DMAX = 10,     // |delta|_inf (bit-pattern max)
    SC_FIXED = 11,       // cost of the residual blocks without a variable parameter block (Ceres' fixed_cost)
    SC_COUNT = 32
};
// PCG scalar slots
enum PcgScal {
    PS_RHO = 0, PS_RHO_NEXT = 1, PS_PQ = 2, PS_Q1 = 3, PS_Q0 = 4, PS_NORMB2 = 5, PS_ALPHA = 6,
    PS_BETA = 7, PS_DONE = 8, PS_ITERS = 9, PS_FAIL = 10, PS_NORMR2 = 11, PS_COUNT = 16
};

// Raw device pointers handed to kernels by value.
struct DevView {
    CameraIntrinsics cam;
    int n_cams, n_free, n_lm;
    long long n_obs;
    const double* poses;      // [n_cams][12] state the pass evaluates at
    const double* points;     // [n_lm][3], internal (landmark-major) order
    const int* cam_free;      // [n_cams] -> free index or -1
    // observation k of landmark j sits at lm_base[j] + k * lm_stride[j]: stride 1 for
    // landmark-major storage, stride G inside a group of G landmarks with identical camera
    // lists (slot-major storage: consecutive landmarks are consecutive in memory)
    const uint32_t* lm_base;
    const uint32_t* lm_stride;
    const uint32_t* lm_cnt;
    const uint32_t* obs_cam;  // [n_obs]
    const double* obs_u;      // SoA observations, internal order
    const double* obs_v;
    const double* obs_d;
    const double* obs_W;      // 9 doubles shared, or 9 per observation (AoS)
    int W_per_obs;
    const double* sc_p;       // [6 n_free] Jacobi column scaling (pose tangent)
    const double* sc_l;       // [3 n_lm]
    // reduced camera system (upper block-CSR, 6x6 row-major blocks)
    const int* s_rowptr;
    const int* s_col;
};

// Landmarks that share one camera list, processed together by the grouped Schur kernel.
// A work item is a slice of at most kItemMax landmarks of one group.
constexpr int kGroupLmax = 16;   // longest camera list the grouped kernel takes
constexpr int kItemMax = 128;    // landmarks per work item
struct GroupView {
    int n_items;
    const int* item_group;       // [n_items]
    const int* item_j0;          // first landmark of the slice, local to the group
    const int* item_n;           // landmarks in the slice
    const int* g_L;              // cameras per landmark
    const int* g_G;              // landmarks in the group
    const int* g_lm0;            // first internal landmark index
    const uint32_t* g_obs0;      // first observation index
    const int* g_off;            // offset of the group's camera list in g_cams
    const int* g_cams;           // camera ids, ascending
    const int* g_blk_off;        // offset of the group's pair table in g_blk
    const int* g_blk;            // S block index for slot pair (a <= b), -1 if a camera is constant
    const int* g_long;           // per group: 1 = accumulate into S_long / Bdiag_long (nullptr: everything into S / Bdiag)
    double *S_long, *Bdiag_long;
    const int* g_map_off;        // ragged groups: offset of [inv: L x G | fwd: K x G] in g_map, -1 for an exact group
    const unsigned char* g_map;  // inv[i][j]: observation index of landmark j in camera slot i (0xff none); fwd[k][j]: slot of observation k
};

// Direct solve of a block-banded reduced camera system (kernels_band.cu).
constexpr int kBandWmax = 12;    // widest half-bandwidth (in 6x6 blocks) the direct solver takes
struct BandView {
    int n, w, P, m;              // block rows, half-bandwidth, leaves, rows per leaf
    int sep_solver;              // separator system: 1 = banded Cholesky on one CTA, 2 = block cyclic reduction
    const int* band_idx;         // [n][w+1] index of block (a, a+d) in the upper block-CSR, or -1
    const double* S;             // block values (diagonal blocks full symmetric)
    const double* rhs;           // [6 n]
    double* Lbuf;                // [n][w+1][36] Cholesky factor, column k: Lkk, L_{k+1,k}, ...
    double* Xbuf;                // [n][6][1+6w] Lkk^-1 [r | B_left] rows
    double *Ta, *Ca, *fa;        // per leaf: separator block, coupling to the previous separator, rhs
    double *Tb, *fb;             // per leaf: X^T X contribution to the separator before it
    double* y;                   // [6 n] solution
    int* fail;
};
// Level-2 (separator) system of the banded solver, dense band storage of width 2W-1.
struct BandScratch {
    double *T2, *rhs2, *L2, *X2, *y2;
};
int band_storage_width(int w);   // 3, 6, 9 or 12: the compiled window widths

// Dense Cholesky of the reduced camera system (kernels_dense.cu): lower triangle, column-major, leading
// dimension ld, n = 6 x free poses padded to n_pad (a multiple of the panel width); row n_pad = right-hand side.
constexpr int kBandPcW = 9;           // half-bandwidth kept by the banded preconditioner (the widest the fast leaf / BCR kernels take)
constexpr int kDenseMaxN = 12288;     // 1.2 GB of FP64
struct DenseView {
    int n, n_pad, ld;
    const int *rowptr, *col;     // upper block-CSR of S
    const double* S;
    const double* rhs;           // [n]
    double* A;                   // [n_pad + 1][ld]
    double* Ldiag;               // [n_pad / 48][48][48]: L11^-T of every diagonal block, [c][i] = (L11^-T)[i][c]
    double* y;                   // [n] solution
    int* fail;
};
int dense_panel_width();

// K3e — exact solve of a WIDE block-banded reduced system (half-bandwidth 13 .. 64 blocks: tracks longer than 13
// frames on a problem too large, or too sparse, for the dense factorisation), kernels_wband.cu.  The poses are cut
// into C chunks with separators of w poses; every chunk is a bordered band (border rows = [left separator | right
// separator | rhs]) factored in band storage with the dense solver's panel / DMMA kernels, all chunks per launch;
// the separator system (block tridiagonal, dense storage) goes through the dense solver.
constexpr int kWbandMaxW = 64;
struct WbandView {
    int n_free, w, C;            // unknown blocks ("units": the free poses at level 0), half-bandwidth (units), chunks
    int unit;                    // scalars per unit: 6 at level 0, 6w of level 0 at level 1 (its separator blocks)
    const WbandView* next;       // host side only: the view that solves this level's separator system (nullptr: dense)
    int m_pad;                   // padded interior size of a chunk (scalars, a multiple of 48; the same for every chunk)
    int sepw;                    // unit * w (0 when C == 1)
    int nbr;                     // border rows: 2 sepw + 1, [left separator | rhs | right separator]
    int r_start;                 // first panel column from which the right separator's rows take part
    int bwr;                     // scalar half-bandwidth unit (w + 1) - 1 rounded up to a multiple of 8
    int ld;                      // band storage: element (i, j), i >= j, of a chunk at A[j * ld + i], ld = bwr + 48
    int ldB;                     // border storage: (border row b, column j) at Bd[j * ldB + b]; columns m_pad + b' hold the
                                 //   border x border Schur complement
    long long a_stride, b_stride;  // doubles per chunk in A / Bd
    const int *rowptr, *col;     // upper block-CSR of S
    const double* S;
    const double* rhs;           // [6 n_free]
    const int* owner;            // per free pose: chunk c >= 0 (interior) or -(s + 1) (separator s)
    const int* local;            // its index inside the chunk interior / the separator
    const int* chunk_p0;         // [C] first pose of the chunk interior
    const int* chunk_len;        // [C] interior poses
    double* A;                   // [C][a_stride]
    double* Bd;                  // [C][b_stride]
    double* Ldiag;               // [C][m_pad / 48][48][48]: L11^-T of every diagonal block, [c][i] = (L11^-T)[i][c]
    double* xw;                  // [C][m_pad]
    double* xsep;                // [(C - 1) sepw] separator unknowns
    DenseView T;                 // separator system (n = (C - 1) sepw; y = xsep)
    double* Txw;                 // [T.n_pad] work vector of the dense solve
    double* y;                   // [6 n_free] solution
    int* fail;
};

struct SunBlockData {
    uint32_t cam;
    double obs_c[3], ref_g[3], W[4], az_thresh, zen_thresh, huber;
};
struct PriorBlockData {
    uint32_t cam;
    double Tref[12], W[36];
};

// Lighting blocks (kernels_phong.cu): raw device pointers, caller's block order.
struct PhongView {
    long long n;
    const uint32_t *cam, *vertex, *material_id;
    const double *intensity, *normal_obs;
    const double *poses, *points, *normals, *texture, *phong, *light, *W_normal;
    const int* cam_free;
    double int_stiffness;
    int directional;
};

// Joint lighting solve (kernels_phong_solve.cu): the state the vertex-elimination kernels read in
// addition to DevView (positions are DevView::points, their scaling DevView::sc_l).
struct PhongSolveView {
    const double* normals;   // [n_lm][3], internal landmark order
    const int* v_mat;        // [n_lm] material of the vertex
    const int* v_tex;        // [n_lm] texture block of the vertex
    const double* gx;        // [n_g] shared blocks: [materials 3 n_mat | textures n_tex | light 3]
    const double* obs_I;     // [n_obs] observed intensity, internal observation order
    const double* obs_n;     // [3][n_obs] observed normal (camera frame), SoA
    const double* sc_n;      // [3 n_lm] column scaling of the normal tangent
    const double* sc_g;      // [n_g]
    const int* g_used;       // [n_g] 1 when some vertex refers to the column
    double Wn[9];
    double int_stiffness;
    int directional, n_mat, n_tex, n_g;
    int hold_positions;      // every position block constant: its columns vanish; a stereo block whose
                             // pose is constant too is dropped (its cost is Ceres' fixed_cost)
    double mat_lo[3], mat_hi[3], tex_lo, tex_hi;
};
// Where the vertex-elimination kernel accumulates the arrowhead reduced system.
struct PhongSystem {
    double *S, *Bdiag, *bp, *gp;  // camera part, as in the stereo path
    double* Scg;                  // [n_g][6 n_free] border, one column after the other
    double* Sgg;                  // [n_g][n_g]
    double* bg;                   // [n_g] reduced right-hand side
    double* gg;                   // [n_g] gradient
    double* hg;                   // [n_g] diag(J_g^T J_g): column norms / LM diagonal
    double* gv;                   // [6 n_lm] vertex gradient [position 3 | normal 3]
    double *cn_l, *cn_n;          // [3 n_lm] each: squared column norms of position / normal (initial pass)
    double* scal;
};

// LM diagonal parameters: D^2 = clamp(diag, min, max) * inv_radius (levenberg_marquardt_strategy)
struct LmDiag {
    double inv_radius, min_diag, max_diag;
};

struct LmRow {
    double v[CSLAM_LOG_COLS];
};

class Engine {
   public:
    explicit Engine(const cslam_options& o);
    ~Engine();

    // ---- host description (borrowed pointers stay owned by the caller) ----
    cslam_options opt;
    CameraIntrinsics cam{1, 1, 0, 0, 1};
    double* h_poses = nullptr;
    uint32_t n_poses = 0;
    std::vector<uint8_t> pose_const;
    double* h_points = nullptr;
    uint32_t n_points = 0;
    uint64_t n_st = 0;
    const uint32_t* st_cam = nullptr;
    const uint32_t* st_pt = nullptr;
    const double* st_uvd = nullptr;
    const double* st_W = nullptr;
    int st_W_per_obs = 0;
    std::vector<SunBlockData> suns;
    // lighting blocks (dataset_ba_phong)
    double* h_normals = nullptr;
    double* h_textures = nullptr;
    const uint32_t* h_material_id = nullptr;
    uint32_t n_vertices = 0;
    double* h_phong = nullptr;
    uint32_t n_materials = 0;
    double* h_tex_shared = nullptr;           // cslam_set_textures: texture blocks shared between vertices
    const uint32_t* h_texture_id = nullptr;
    uint32_t n_tex_shared = 0;
    double mat_lo[3] = {-1e308, -1e308, -1e308}, mat_hi[3] = {1e308, 1e308, 1e308};
    double tex_lo = -1e308, tex_hi = 1e308;
    bool bounded = false;                     // cslam_set_bounds was called
    bool hold_positions = false;              // cslam_set_points_constant
    double* h_light = nullptr;
    int light_directional = 0;
    uint64_t n_ph = 0;
    const uint32_t* ph_cam = nullptr;
    const uint32_t* ph_vertex = nullptr;
    const double* ph_intensity = nullptr;
    const double* ph_normal_obs = nullptr;
    double ph_int_stiffness = 1.0;
    double ph_W_normal[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    std::vector<PriorBlockData> priors;
    std::string err;

    // multi-GPU
    int n_ranks = 1, rank = 0;
    void* nccl_comm = nullptr;

    // ---- entry points used by the C ABI ----
    void upload();
    void lm_begin();
    void lm_iterate(int n, bool ignore_convergence, cslam_summary* s);
    void download();
    void reset_state();
    void evaluate(int apply_loss, double* cost, double* r_st, double* Jc_st, double* Jp_st, double* r_sun,
                  double* J_sun, double* r_pr, double* J_pr);
    double time_resjac(int reps);
    void covariance_block(uint32_t cam, double* cov36);
    void evaluate_phong(double* cost, double* r_int, double* J_int, double* r_n, double* Jc_n, double* Jn_n);
    double time_phong(int reps);
    double time_schur(int reps);
    void fill_summary(cslam_summary* s) const;
    void get_reduced_sizes(int* nf, int* nnz) const;
    void get_reduced_system(int* rowptr, int* col, double* values, double* rhs, int* ids);
    void set_stream(cudaStream_t s);
    void analyze(int n_ranks_, int rank_, cslam_structure_info* out);  // host only

    double fixed_cost() const { return lm.fixed_cost; }
    std::vector<LmRow> log;
    cslam_profile prof{};
    bool uploaded = false, begun = false;
    bool phong_ready = false;             // block-index arrays of the lighting blocks are on the device
    bool lighting_in_solve() const { return n_ph > 0; }

    // small-problem description used by the batched window kernel
    bool window_eligible(bool any_strategy = false) const;
    void set_window_summary(const cslam_summary& s);

   private:
    friend struct EngineAccess;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr;

    // structure (host copies kept for download / diagnostics)
    std::vector<int> cam_free_h, free_cams_h;
    std::vector<uint32_t> lm_user_h;      // internal landmark -> user point index
    std::vector<uint32_t> lm_base_h, lm_stride_h, lm_cnt_h;
    int n_lm_grouped = 0;                 // internal landmarks [0, n_lm_grouped) belong to groups
    // internal obs -> user obs index; not value-initialised (80 MB for config 5: the pages are first
    // touched by the parallel fill, not by one thread zeroing them)
    std::unique_ptr<uint32_t[]> obs_user_h;
    size_t obs_user_n = 0;
    std::vector<int> s_rowptr_h, s_col_h;
    int n_free = 0, n_lm = 0;
    long long n_obs = 0;
    int nnzU = 0;
    // shard of landmarks this rank owns (internal order)
    int lm_lo = 0, lm_hi = 0;

    // device state
    DBuf<double> d_poses, d_poses_cand, d_poses_best, d_poses_init;
    DBuf<double> d_points, d_points_cand, d_points_best, d_points_init;
    DBuf<int> d_cam_free;
    DBuf<uint32_t> d_lm_base, d_lm_stride, d_lm_cnt, d_obs_cam;
    // caller-order staging on the device (upload gathers from it, download scatters into d_raw_pts)
    DBuf<uint32_t> d_raw_cam, d_raw_pt, d_obs_user, d_lm_user;
    bool structure_on_device = false;  // layout tables and the observation permutation were built on the device
    DBuf<double> d_raw_uvd, d_raw_W, d_raw_pts;
    // grouped Schur path
    std::vector<int> item_group_h, item_j0_h, item_n_h, g_L_h, g_G_h, g_lm0_h, g_off_h, g_cams_h, g_blk_off_h, g_blk_h;
    // ragged groups (landmarks whose cameras fit a common window of <= 10 without sharing the exact list)
    std::vector<int> g_map_off_h;
    std::vector<unsigned char> g_map_h;
    int n_items_rag = 0;
    bool want_ragged = true;       // false while a host analysis is only rebuilt to verify the device one
    long long n_obs_true = 0;      // observations without the padding rows of ragged groups
    DBuf<int> d_g_map_off;
    DBuf<unsigned char> d_g_map;
    std::vector<uint32_t> g_obs0_h;
    DBuf<int> d_item_group, d_item_j0, d_item_n, d_g_L, d_g_G, d_g_lm0, d_g_off, d_g_cams, d_g_blk_off, d_g_blk;
    DBuf<uint32_t> d_g_obs0;
    int max_group_L = 0, n_items_small = 0;
    GroupView group_view() const;
    void launch_schur(const DevView& v, const LmDiag& dg);
    // wide-window slices (K2w): runs of ungrouped landmarks whose cameras fit [c0, c0 + 32); host analysis only
    std::vector<int> wide_lo_h, wide_hi_h, wide_c0_h;
    std::vector<uint8_t> wide_flag_h;      // per ungrouped landmark: 1 = in a slice
    long long wide_obs0 = 0;               // first observation of the ungrouped landmarks
    long long n_wide_lm = 0;
    DBuf<int> d_wide_lo, d_wide_hi, d_wide_c0;
    DBuf<uint8_t> d_wide_flag;
    DBuf<double> d_wide_Z;                 // [observations of the ungrouped landmarks][18]
    DBuf<double> d_obs_u, d_obs_v, d_obs_d, d_obs_W;
    DBuf<double> d_sc_p, d_sc_l, d_cn_p, d_cn_l;
    DBuf<double> d_gl;                     // scaled point gradient from the last Schur pass
    DBuf<int> d_s_rowptr, d_s_col, d_lt_rowptr, d_lt_col, d_lt_blk;
    // [ S values (36 nnzU) | Bdiag (36 nf) | rhs (6 nf) | gp (6 nf) | scalars (SC_COUNT) ] — one
    // contiguous buffer so a single all-reduce sums every rank's partial reduced system
    DBuf<double> d_red;
    double *d_S = nullptr, *d_Bdiag = nullptr, *d_bp = nullptr, *d_gp = nullptr, *d_scal = nullptr;
    size_t red_count = 0;
    DBuf<double> d_Minv, d_diag_p;
    DBuf<double> d_yp, d_pr, d_pz, d_pp, d_pq, d_yl, d_pp2, d_prec;
    DBuf<double> d_pscal;
    // banded direct solver (linear_solver == 0 and S block-banded)
    int band_w = 0, band_P = 0, band_m = 0, band_sep = 1;
    bool band_active = false;
    DBuf<int> d_band_idx, d_band_fail;
    DBuf<double> d_Lbuf, d_Xbuf, d_Ta, d_Ca, d_fa, d_Tb, d_fb, d_T2, d_rhs2, d_L2, d_X2, d_y2;
    void plan_band_solver();
    // CG preconditioned with the banded solver (S neither a narrow band nor small enough for the dense factorisation)
    bool bandpc_active = false;
    DBuf<double> d_bpc_S, d_bpc_scal, d_S2, d_Bdiag2;   // S2 / Bdiag2: contributions of the landmarks beyond the window
    DBuf<int> d_g_long;
    void bandpc_solve(const double* rhs, double* y);
    // dense direct solver (linear_solver == 0, S not banded, small or dense enough)
    bool dense_active = false;
    int dense_npad = 0, dense_ld = 0;
    DBuf<double> d_dense_A, d_dense_Ld, d_dense_xw;
    void plan_dense_solver();
    // wide-band direct solver (linear_solver == 0, S block-banded with 12 < half-bandwidth <= 64)
    bool wband_active = false;
    // level 0: the poses (unit 6, half-bandwidth w); level 1, when there are enough separators: the separator system of
    // level 0, block tridiagonal with blocks of 6w — the same chunking again (unit 6w, half-bandwidth 1 block)
    struct WbandLevel {
        int n_units = 0, unit = 0, w = 0, C = 0, m_pad = 0, r_start = 0;
        DBuf<int> owner, local, p0, len;
        DBuf<double> A, Bd, Ld, xw, xsep, T, TLd, Txw;
    };
    WbandLevel wb[2];
    int wb_levels = 0;
    bool plan_wband_level(int lvl, int n_units, int unit, int w, int C);
    WbandView wband_view(int lvl);
    bool plan_wband_solver(int w);
    // extra scratch sets + streams so that independent solves against the same banded S run
    // concurrently (the border columns of the lighting solve): each solve is a latency-bound chain
    // on a handful of CTAs, so n_g + 1 of them fit side by side on 148 SMs
    struct BandSet {
        int* fail = nullptr;       // all pointers are slices of band_pool
        double *Lbuf = nullptr, *Xbuf = nullptr, *Ta = nullptr, *Ca = nullptr, *fa = nullptr, *Tb = nullptr, *fb = nullptr,
               *T2 = nullptr, *rhs2 = nullptr, *L2 = nullptr, *X2 = nullptr, *y2 = nullptr, *ps = nullptr;
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        ~BandSet() {
            if (done) cudaEventDestroy(done);
            if (stream) cudaStreamDestroy(stream);
        }
    };
    std::vector<std::unique_ptr<BandSet>> band_sets;
    DBuf<double> band_pool;
    cudaEvent_t ev_fork = nullptr;
    void alloc_band_sets(int count);
    void solve_reduced_on(BandSet& bs, const double* rhs, double* y);
    DBuf<double> d_scal2;                  // scalars of the back-substitution pass
    DBuf<SunBlockData> d_suns;
    DBuf<PriorBlockData> d_priors;
    // user-order copy for the materialised residual/Jacobian kernel
    DBuf<uint32_t> d_u_cam, d_u_pt;
    DBuf<double> d_u_u, d_u_v, d_u_d, d_u_W, d_u_points;
    DBuf<int> d_u_tile_lo, d_u_tile_n;
    DBuf<double> d_o_r, d_o_Jc, d_o_Jp;
    bool user_copy_ready = false;
    // lighting blocks, caller's order
    DBuf<uint32_t> d_ph_cam, d_ph_vertex, d_ph_mat;
    DBuf<double> d_ph_int, d_ph_nobs, d_ph_normals, d_ph_tex, d_ph_phong, d_ph_light, d_ph_W, d_ph_points, d_ph_poses;
    DBuf<double> d_ph_rI, d_ph_JI, d_ph_rN, d_ph_JNc, d_ph_JNn;
    DBuf<int> d_ph_cam_free;
    PhongView phong_view();
    void ensure_phong();
    double* h_pinned = nullptr;            // pinned scalar read-back
    // joint lighting solve (dataset_ba_phong stage 3): vertex = position + normal, shared blocks gx
    struct PhongSolve {
        bool active = false;
        int n_g = 0, n_mat = 0, n_tex = 0, max_track = 0;
        DBuf<double> normals, normals_cand, normals_best, normals_init;
        DBuf<double> gx, gx_cand, gx_best, gx_init;
        DBuf<int> v_mat, v_tex, g_used;
        DBuf<double> obs_I, obs_n;
        DBuf<double> sc_n, sc_g, cn_n, gv, yv, yg, diag_g, X, T, zero_g;
        DBuf<double> diag_v, sc_v, Yv, Yg;      // DOGLEG: clamp(diag(J^T J)) and column scaling per vertex, combined step
        double *Scg = nullptr, *Sgg = nullptr, *bg = nullptr, *gg = nullptr, *hg = nullptr;  // inside d_red
        std::vector<int> g_used_h;
        cudaGraphExec_t fan_graph = nullptr;   // the border solves' fan-out, captured on its second use
        unsigned long long fan_graph_kernels = 0;
        int fan_calls = 0;
    } ph;
    void check_phong_solve();
    void setup_phong_solve();
    void copy_phong_best();
    void phong_step(const LmDiag& dg, double* sc2);
    void phong_line_search(const double* yp, const double* yg, const double* yv, double* sc2);
    void phong_reduce_scal2();
    DBuf<double> d_dmax_tmp;
    void phong_dogleg_step(int* lin_iters, bool* valid, double* sc2);
    PhongSolveView phong_solve_view(const double* normals, const double* gx) const;
    PhongSystem phong_system();
    void solve_reduced(const double* rhs, double* y);
    void phong_linear_solve(int* iters, bool* ok);

    // LM state (mirrors oracle/problem.hpp::solve)
    struct Lm {
        double x_cost = 0, x_norm = 0, gradient_max_norm = 0;
        double radius = 0, decrease_factor = 2;
        bool reuse_diagonal = false;
        double minimum_cost = 0;
        double se_minimum = 0, se_current = 0, se_reference = 0, se_candidate = 0, se_acc_ref = 0,
               se_acc_cand = 0;
        int se_nonmono = 0;
        int invalid_steps = 0;
        int iteration = 0;
        bool step_ok_prev = false, finished = false;
        bool have_system = false;          // S, rhs valid for current (x, radius)
        bool grad_fresh = false;
        int num_successful = 0, num_unsuccessful = 0, total_linear = 0;
        int termination_type = 1, termination_reason = 0;
        double initial_cost = 0;
        double fixed_cost = 0;             // reported with every cost, no part in the decisions
        double device_ms = 0;
    } lm;

    // DOGLEG trust-region strategy (opt.trust_region_strategy == 1), stereo / sun / prior problems
    struct Dogleg {
        double mu = 1e-8, step_norm = 0;
        bool reuse = false;
        DoglegModel model;
    } dl;
    DBuf<double> d_diag_l, d_Yp, d_Yl, d_dsums;
    bool dogleg() const { return opt.trust_region_strategy == 1; }
    LmDiag current_diag() const {
        return LmDiag{dogleg() ? dl.mu : 1.0 / lm.radius, opt.min_lm_diagonal, opt.max_lm_diagonal};
    }
    void dogleg_step(int* lin_iters, bool* valid, double* sc2);

    DevView view(const double* poses, const double* points) const;
    void build_structure();
    bool build_structure_gpu(cudaEvent_t idx_ready);  // false: fall back to the host analysis
    unsigned long long layout_hash() const;
    void ensure_user_copy();
    void schur_pass();
    void run_pcg(int* iters, bool* ok);
    void gradient_norm_pass();
    void prof_begin(int k);
    void prof_end(int k);
    void read_scalars(const double* dev, double* host, int n);
    void allreduce_system();
    void allreduce_small(double* dev, int n);
};

// window batch (kernels_window.cu)
// cov_cam >= 0 (n == 1): instead of solving, the marginal covariance block of that pose at the caller's current values
// (cslam_covariance_block for a window-eligible problem: one launch of window_cov_kernel), 36 doubles to cov_out
void solve_window_batch(Engine** engines, int n, cslam_summary* summaries, int cov_cam = -1, double* cov_out = nullptr);

}  // namespace cslam
