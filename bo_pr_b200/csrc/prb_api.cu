// C ABI of libprb200 (see include/prb200.h).  Host-side orchestration only: argument validation, workspace carving and
// kernel sequencing on the caller's stream.  No allocation and no synchronisation on the hot entry points.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "prb_device.cuh"
#include "prb_kernels.cuh"

using namespace prb;

#include <mutex>
#include <vector>

namespace prb {
std::atomic<unsigned long long> g_launches{0};
thread_local bool g_pdl_on = false;  // per host thread: a PdlScope only affects the launches of the call that opened it

const EnvKnobs& env_knobs() {
  static const EnvKnobs k = [] {
    EnvKnobs e;
    e.no_pdl = getenv("PRB_NO_PDL") != nullptr;
    e.no_fused_step = getenv("PRB_NO_FUSED_STEP") != nullptr;
    e.panel_mb = getenv("PRB_PANEL_MB") ? atol(getenv("PRB_PANEL_MB")) : 0;
    e.force_bn = getenv("PRB_BN") ? atoi(getenv("PRB_BN")) : 0;
    e.group_waves = getenv("PRB_GROUP_WAVES") ? atoi(getenv("PRB_GROUP_WAVES")) : 4;
    e.mixed_order = getenv("PRB_MIXED_ORDER") ? atoi(getenv("PRB_MIXED_ORDER")) : 1;
    return e;
  }();
  return k;
}

static std::atomic<bool> g_prof_on{false};
struct ProfRec { int id; cudaEvent_t beg, end; };
static std::vector<ProfRec> g_prof;  // diagnostic record (prb_profile_*), guarded by g_prof_mu
static std::mutex g_prof_mu;

ProfScope::ProfScope(int id, cudaStream_t st) : id_(id), st_(st), beg_(nullptr), end_(nullptr), on_(g_prof_on) {
  if (!on_) return;
  if (cudaEventCreate(&beg_) != cudaSuccess || cudaEventCreate(&end_) != cudaSuccess) {
    on_ = false;
    return;
  }
  cudaEventRecord(beg_, st_);
}
ProfScope::~ProfScope() {
  if (!on_) return;
  cudaEventRecord(end_, st_);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back({id_, beg_, end_});
}
}  // namespace prb

namespace {

thread_local char g_cuda_err[256] = "";

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
  return PRB_ERR_CUDA;
}
#define CUDA_TRY(expr)                                  \
  do {                                                  \
    cudaError_t _e = (expr);                            \
    if (_e != cudaSuccess) return cuda_fail(_e, #expr); \
  } while (0)

// ---- GP-state packing --------------------------------------------------------------------------------------------
// A1[j, k] = Linv[j, k] (k <= j < n) ; A1[n, k] = alpha[k] ; 0 elsewhere.      [Mp1, Kp]
// A2[m, k] = Linv[k, m] (m <= k < n) ; 0 elsewhere.                             [Mp2, Kp]
__global__ void pack_state_kernel(const double* __restrict__ Linv, const double* __restrict__ alpha, int n, int Kp,
                                  int Mp1, int Mp2, double* __restrict__ A1, double* __restrict__ A2) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (k >= Kp) return;
  if (r < Mp1) {
    double v = 0.0;
    if (r < n && k <= r) v = Linv[size_t(r) * n + k];
    if (r == n && k < n) v = alpha[k];
    A1[size_t(r) * Kp + k] = v;
  }
  if (r < Mp2) {
    double v = 0.0;
    if (r < n && k < n && k >= r) v = Linv[size_t(k) * n + r];
    A2[size_t(r) * Kp + k] = v;
  }
}

int make_spec(const prb_model_desc* d, DevSpec* out, double* kss) {
  if (d->n_terms < 1 || d->n_terms > PRB_MAX_TERMS) return PRB_ERR_UNSUPPORTED;
  memset(out, 0, sizeof(DevSpec));
  out->n_terms = d->n_terms;
  out->d_k = d->d_k;
  double diag = 0.0;
  for (int t = 0; t < d->n_terms; ++t) {
    const prb_term& T = d->terms[t];
    if (T.n_factors < 1 || T.n_factors > PRB_MAX_FACTORS) return PRB_ERR_UNSUPPORTED;
    DevTerm& D = out->t[t];
    D.outputscale = T.outputscale;
    D.additive = T.additive ? 1 : 0;
    D.n_factors = T.n_factors;
    for (int f = 0; f < T.n_factors; ++f) {
      const prb_factor& F = T.factors[f];
      if (F.n_dims < 1 || F.n_dims > PRB_MAX_DIMS || F.power < 1 || F.power > 4) return PRB_ERR_UNSUPPORTED;
      if (F.kind != PRB_FACTOR_MATERN52 && F.kind != PRB_FACTOR_CATEGORICAL) return PRB_ERR_INVALID;
      DevFactor& G = D.f[f];
      G.kind = F.kind;
      G.n_dims = F.n_dims;
      G.power = F.power;
      for (int i = 0; i < F.n_dims; ++i) {
        if (F.dims[i] < 0 || F.dims[i] >= d->d_k || !(F.lengthscale[i] > 0.0)) return PRB_ERR_INVALID;
        G.dims[i] = F.dims[i];
        G.inv_ls[i] = (F.kind == PRB_FACTOR_MATERN52) ? 1.0 / F.lengthscale[i] : 1.0 / (F.lengthscale[i] * F.n_dims);
      }
    }
    diag += T.outputscale * (T.additive ? double(T.n_factors) : 1.0);
  }
  int nu = 0;
  bool fits = true;
  for (int t = 0; t < d->n_terms && fits; ++t)
    for (int f = 0; f < out->t[t].n_factors && fits; ++f) {
      const DevFactor& F = out->t[t].f[f];
      int u = -1;
      for (int k = 0; k < nu && u < 0; ++k) {
        const DevFactor& G = out->t[out->first_t[k]].f[out->first_f[k]];
        bool same = G.kind == F.kind && G.n_dims == F.n_dims;
        for (int i = 0; i < F.n_dims && same; ++i) same = G.dims[i] == F.dims[i] && G.inv_ls[i] == F.inv_ls[i];
        if (same) u = k;
      }
      if (u < 0) {
        if (nu == 4) {
          fits = false;
          break;
        }
        u = nu++;
        out->first_t[u] = int8_t(t);
        out->first_f[u] = int8_t(f);
      }
      out->uid[t][f] = int8_t(u);
    }
  out->n_unique = fits ? int8_t(nu) : int8_t(0);
  *kss = diag;
  return PRB_OK;
}

int make_space(const prb_space_desc* s, DevSpace* out) {
  if (!s || s->d < 1 || s->d > PRB_MAX_DIMS || s->n_int < 0 || s->n_cat < 0 || s->n_cat > PRB_MAX_CAT)
    return PRB_ERR_UNSUPPORTED;
  if (s->n_int + s->n_cat == 0 || !(s->tau > 0.0)) return PRB_ERR_INVALID;
  memset(out, 0, sizeof(DevSpace));
  out->d = s->d;
  out->n_int = s->n_int;
  out->n_cat = s->n_cat;
  out->apply_numeric = s->apply_numeric ? 1 : 0;
  out->tau = s->tau;
  out->raw_int = s->raw_integer_space ? 1 : 0;
  bool is_disc[PRB_MAX_DIMS] = {false};
  for (int k = 0; k < s->n_int; ++k) {
    const int c = s->int_cols[k];
    if (c < 0 || c >= s->d || !(s->int_hi[k] > s->int_lo[k])) return PRB_ERR_INVALID;
    out->int_cols[k] = c;
    out->int_lo[k] = s->int_lo[k];
    out->int_hi[k] = s->int_hi[k];
    is_disc[c] = true;
  }
  int n_disc = s->n_int;
  int expect = s->n_cat > 0 ? s->cat_start[0] : s->d;
  for (int j = 0; j < s->n_cat; ++j) {
    if (s->cat_start[j] != expect || s->cat_card[j] < 1 || s->cat_card[j] > PRB_MAX_CARD) return PRB_ERR_INVALID;
    out->cat_start[j] = s->cat_start[j];
    out->cat_card[j] = s->cat_card[j];
    for (int k = 0; k < s->cat_card[j]; ++k) is_disc[expect + k] = true;
    expect += s->cat_card[j];
    n_disc += s->cat_card[j];
  }
  if (expect != s->d) return PRB_ERR_INVALID;  // one-hot blocks must run contiguously to the last column
  out->n_disc = n_disc;
  out->d_k = (s->apply_numeric && s->n_cat > 0) ? s->cat_start[0] + s->n_cat : s->d;
  if (s->latent_tables != nullptr && s->n_cat > 0) {
    if (!s->apply_numeric) return PRB_ERR_INVALID;
    int col = 0, off = 0;
    for (int j = 0; j < s->n_cat; ++j) {
      if (s->latent_dim[j] < 1 || s->latent_dim[j] > PRB_MAX_DIMS) return PRB_ERR_INVALID;
      out->latent_dim[j] = s->latent_dim[j];
      out->latent_col[j] = col;
      out->latent_off[j] = off;
      col += s->latent_dim[j];
      off += s->latent_dim[j] * s->cat_card[j];
    }
    out->latent_tab = s->latent_tables;
    out->d_k = s->cat_start[0] + col;
    if (out->d_k > PRB_MAX_DIMS) return PRB_ERR_UNSUPPORTED;
  }
  int nc = 0;
  for (int c = 0; c < s->d; ++c)
    if (!is_disc[c]) out->cont_cols[nc++] = c;
  out->n_cont = nc;
  return PRB_OK;
}

int make_acq(const prb_acq_desc* a, DevAcq* out) {
  if (!a || a->kind < PRB_ACQ_EI || a->kind > PRB_ACQ_MEAN) return PRB_ERR_INVALID;
  out->kind = a->kind;
  out->ei_variant = a->ei_variant;
  out->param = a->param;
  out->min_variance = a->min_variance;
  out->acq_var_floor = a->acq_var_floor;
  return PRB_OK;
}

// ---- workspace plan ----------------------------------------------------------------------------------------------
struct Plan {
  int64_t B, B_pad;
  int C, nchunks;
  size_t off_xk, off_z, off_prob, off_a, off_dmu, off_dvar, off_S_all, off_probs;
  size_t off_panelA, off_panelV, off_panelQ, off_norm, off_mu, off_S_part;
  size_t off_Xc, off_grad, off_ones, off_val, off_step;
  size_t off_zbits, off_kc, off_Dc;
  size_t off_done, off_small, small_bytes;
  int grad_jrows;  // training rows per CTA of the generic gradient contraction (16 for small panels, else 128)
  size_t off_ucount, off_ustart, off_utotal, off_ucode, off_uweight, off_inv, off_zbits_u, off_colb, off_wu;
  size_t total;
};

// Column panels (k*, v, w) are sized to stay resident in the 126 MB L2 between the kernel that writes them and the one
// that reads them.  PRB_PANEL_MB overrides the budget (tuning knob, read once).
size_t panel_budget_bytes() {
  const long mb = env_knobs().panel_mb;
  return size_t(mb > 0 ? mb : 96) << 20;
}

Plan make_plan(const Model& m, int64_t n_points, int n_restarts, int d_theta, int chunk_cols) {
  Plan p;
  memset(&p, 0, sizeof(p));
  p.B = n_points;
  p.B_pad = round_up64(n_points > 0 ? n_points : 1, 128);
  int64_t C = chunk_cols;
  if (C <= 0) {
    // Generic path: k* and v panels share the budget.  Separable path: only the v panel exists and it is streamed once
    // each way (2 x 8 n bytes per column against 2 n^2 flop), so L2 residency does not matter and wide panels win:
    // fewer launches, and the partial last wave of the triangular tile schedule is amortised over more tiles
    // (measured: 96 MB -> 5.22 ms, 1.1 GB -> 4.77 ms per 65 536-column step at n = 1000).
    // PR calls (n_restarts > 0): panels wide enough that the benchmark shape (65 536 columns at n = 1000) is ONE panel and
    // runs the fused step; the k* / v / Q panels are each streamed once per product, so L2 residency does not matter.
    size_t budget = panel_budget_bytes();
    if (n_restarts > 0 && env_knobs().panel_mb <= 0) budget = size_t(2048) << 20;
    C = int64_t(budget / (sizeof(double) * size_t(m.Mp1 + m.Mp2 + (m.sep_ok && n_restarts > 0 ? 0 : m.Kp)))) / 128 * 128;
    if (C < 128) C = 128;
  }
  C = round_up64(C, 128);
  if (C > p.B_pad) C = p.B_pad;
  int64_t nch = (p.B_pad + C - 1) / C;
  C = round_up64((p.B_pad + nch - 1) / nch, 128);
  p.C = int(C);
  p.nchunks = int((p.B_pad + C - 1) / C);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) / 256 * 256;
    return o;
  };
  const size_t dbl = sizeof(double);
  p.off_xk = take(size_t(p.B_pad) * m.d_k * dbl);
  p.off_z = take(size_t(p.B_pad) * d_theta);
  p.off_prob = take(size_t(p.B_pad) * d_theta * dbl);
  p.off_a = take(size_t(p.B_pad) * dbl);
  p.off_dmu = take(size_t(p.B_pad) * dbl);
  p.off_dvar = take(size_t(p.B_pad) * dbl);
  p.off_S_all = take(size_t(p.B_pad) * m.d_k * dbl);
  p.off_probs = take(size_t(p.B_pad) * dbl);
  p.off_panelA = take(size_t(m.Mp2) * C * dbl);
  p.off_panelV = take(size_t(C) * m.Mp1 * dbl);
  p.off_panelQ = take(size_t(C) * m.Kp * dbl);  // derivative-factor panel of the generic path (kstar_q_kernel)
  p.off_norm = take(size_t(m.Mp1 / GEMM_BM) * C * dbl);
  p.off_mu = take(size_t(C) * dbl);
  // small panels: 4 training rows per CTA of the generic gradient contraction instead of 128 (32x the CTAs; the kernel is
  // a serial loop over its rows: 90 us at n = 119 with 128 rows per CTA, 36 us with 16)
  p.grad_jrows = ((m.Mp2 / GEMM_BM) * (C / 128) < 120) ? 4 : GRAD_JSPLIT_ROWS;
  p.off_S_part = take(size_t(m.Mp2 / p.grad_jrows) * m.d_k * C * dbl);
  p.off_Xc = take(size_t(p.B_pad) * d_theta * dbl);
  p.off_grad = take(size_t(p.B_pad) * d_theta * dbl);
  p.off_ones = take(size_t(p.B_pad) * dbl);
  p.off_val = take(size_t(p.B_pad) * dbl);
  p.off_step = take(256);
  p.off_zbits = take(size_t(p.B_pad) * sizeof(uint32_t));
  p.off_done = take(256);
  if (n_restarts > 0) {  // scratch of the single-kernel step (prb_small.cu)
    p.small_bytes = (size_t(p.B_pad / 32 + n_restarts) * (1 + d_theta)) * dbl + size_t(n_restarts + 4) * sizeof(int) + 64;
    p.off_small = take(p.small_bytes);
  }
  if (m.sep_ok && n_restarts > 0) {
    p.off_kc = take(size_t(n_restarts) * m.Kp * dbl);
    p.off_Dc = take(size_t(n_restarts) * d_theta * m.Kp * dbl);
    p.off_ucount = take(size_t(n_restarts) * sizeof(int));
    p.off_ustart = take(size_t(n_restarts + 1) * sizeof(int));
    p.off_utotal = take(2 * sizeof(int));
    p.off_ucode = take(size_t(p.B_pad) * sizeof(uint32_t));
    p.off_uweight = take(size_t(p.B_pad) * sizeof(int));
    p.off_inv = take(size_t(p.B_pad) * sizeof(int));
    p.off_zbits_u = take(size_t(p.B_pad) * sizeof(uint32_t));
    p.off_colb = take(size_t(p.B_pad) * sizeof(int));
    p.off_wu = take(size_t(p.B_pad) * dbl);
  }
  p.total = off;
  return p;
}

template <typename T>
T* at(void* ws, size_t off) {
  return reinterpret_cast<T*>(static_cast<char*>(ws) + off);
}

// ---- the posterior/acquisition engine over L2-resident column panels ------------------------------------------------
// generic kernel structures: materialised k* panel -> lower product -> epilogue -> upper product -> gradient contraction
int run_engine(const Model& m, const DevAcq& aq, const Plan& pl, void* ws, const double* xk, int64_t B, bool need_grad,
               uint64_t diffmask, double* a, double* dmu, double* dvar, double* S_all, double* out_mean,
               double* out_var, cudaStream_t st) {
  double* panelA = at<double>(ws, pl.off_panelA);
  double* panelV = at<double>(ws, pl.off_panelV);
  double* norm = at<double>(ws, pl.off_norm);
  double* mu = at<double>(ws, pl.off_mu);
  double* S_part = at<double>(ws, pl.off_S_part);
  double* panelQ = at<double>(ws, pl.off_panelQ);
  GradGroup gq;
  const bool gq_ok = need_grad && m.rff_m == 0 && build_grad_group(m.spec, diffmask, &gq);

  if (m.rff_m > 0) {  // deterministic posterior sample: the acquisition value is the sample path itself (PosteriorMean)
    if (aq.kind != PRB_ACQ_MEAN) return PRB_ERR_INVALID;
    CUDA_TRY(launch_rff_eval(m, xk, B, diffmask, need_grad, a, S_all, out_mean, out_var, st));
    return PRB_OK;
  }
  for (int64_t c0 = 0; c0 < B; c0 += pl.C) {
    const int ncols = int(B - c0 < pl.C ? B - c0 : pl.C);
    const int ncols_pad = round_up(ncols, 128);
    const double* xkc = xk + size_t(c0) * m.d_k;
    CUDA_TRY(launch_kstar_panel(m, xkc, ncols, ncols_pad, panelA, st, gq_ok ? &gq : nullptr, panelQ));
    FusedAcq fa;  // single row tile: the acquisition epilogue runs inside the lower product
    fa.aq = aq;
    fa.a = a + c0;
    fa.dmu = dmu + c0;
    fa.dvar = dvar + c0;
    fa.ncols = ncols;
    const bool fuse = m.Mp1 == GEMM_BM && out_mean == nullptr && out_var == nullptr;
    CUDA_TRY(launch_tma_lower(m, panelA, pl.C, ncols_pad, need_grad ? panelV : nullptr, norm, mu, fuse ? &fa : nullptr,
                              st));
    if (!fuse)
      CUDA_TRY(launch_acq_epilogue(m, aq, norm, ncols_pad, mu, ncols, a + c0, dmu + c0, dvar + c0,
                                   out_mean ? out_mean + c0 : nullptr, out_var ? out_var + c0 : nullptr, nullptr, 0,
                                   st));
    if (need_grad && gq_ok) {
      // gradient contraction fused into the upper product's epilogue: no W panel, no second kernel evaluation
      CUDA_TRY(launch_tma_upper_gradq(m, panelV, pl.C, ncols, ncols_pad, gq, panelQ, xkc, dmu + c0, dvar + c0, S_part, st));
      CUDA_TRY(launch_sum_splits(m, S_part, m.Mp2 / GEMM_BM, nullptr, ncols, ncols_pad, S_all + size_t(c0) * m.d_k,
                                 nullptr, 0, st));
    } else if (need_grad) {
      // several distinct Matern factors carry a gradient (e.g. d acq / dx over the binary factor's dims too): W panel +
      // the general contraction kernel
      CUDA_TRY(launch_tma_upper(m, panelV, pl.C, ncols_pad, panelA, st));
      CUDA_TRY(launch_grad_contract(m, diffmask, xkc, ncols, ncols_pad, panelA, dmu + c0, dvar + c0, S_part,
                                    pl.grad_jrows, st));
      CUDA_TRY(launch_sum_splits(m, S_part, m.Mp2 / pl.grad_jrows, nullptr, ncols, ncols_pad,
                                 S_all + size_t(c0) * m.d_k, nullptr, 0, st));
    }
  }
  return PRB_OK;
}

// Does this PR space line up with the model's separable structure?  Every integer parameter must be a binary dim of the
// binary factor (bounds [0, 1]), there are no categoricals, and the continuous columns are exactly the other factor's.
bool sep_space_ok(const Model& m, const DevSpace& sp, BitMap* bm, DimMap* dm) {
  if (!m.sep_ok || sp.n_cat != 0 || sp.d != m.d_k || sp.raw_int) return false;
  const DevFactor& fb = m.spec.t[0].f[m.sep_bin_factor];
  const DevFactor& fc = m.spec.t[0].f[m.sep_cont_factor];
  if (sp.n_int != fb.n_dims || sp.n_cont != fc.n_dims) return false;
  for (int k = 0; k < PRB_MAX_DIMS; ++k) bm->pos[k] = -1;
  for (int k = 0; k < sp.n_int; ++k) {
    if (sp.int_lo[k] != 0.0 || sp.int_hi[k] != 1.0) return false;
    int pos = -1;
    for (int di = 0; di < fb.n_dims; ++di)
      if (fb.dims[di] == sp.int_cols[k]) pos = di;
    if (pos < 0) return false;
    bm->pos[k] = pos;
  }
  dm->n = fc.n_dims;
  for (int di = 0; di < fc.n_dims; ++di) {
    bool found = false;
    for (int ci = 0; ci < sp.n_cont; ++ci) found = found || sp.cont_cols[ci] == fc.dims[di];
    if (!found) return false;
    dm->dims[di] = fc.dims[di];
  }
  return true;
}

// separable fast path: no k* panel, no W panel -- operands generated in the lower product, gradient contracted in the
// epilogue of the upper product
int run_engine_sep(const Model& m, const DevAcq& aq, const Plan& pl, void* ws, const double* theta, int d_theta, int nb,
                   int mc, const DimMap& dm, bool need_grad, bool dedup, double* a, double* dmu, double* dvar,
                   double* S_all, cudaStream_t st) {
  const int64_t B = int64_t(nb) * mc;  // with de-duplication: the worst case (no two samples of a restart coincide)
  double* panelV = at<double>(ws, pl.off_panelV);
  double* norm = at<double>(ws, pl.off_norm);
  double* mu = at<double>(ws, pl.off_mu);
  double* S_part = at<double>(ws, pl.off_S_part);
  double* kc = at<double>(ws, pl.off_kc);
  double* Dc = at<double>(ws, pl.off_Dc);
  const uint32_t* zbits = dedup ? at<uint32_t>(ws, pl.off_zbits_u) : at<uint32_t>(ws, pl.off_zbits);
  const int* colb = dedup ? at<int>(ws, pl.off_colb) : nullptr;
  const int* ncols_dev = dedup ? at<int>(ws, pl.off_utotal) : nullptr;
  CUDA_TRY(launch_sep_restart_prep(m, theta, nb, d_theta, kc, need_grad ? Dc : nullptr, st));
  for (int64_t c0 = 0; c0 < B; c0 += pl.C) {
    const int ncols = int(B - c0 < pl.C ? B - c0 : pl.C);
    const int ncols_pad = round_up(ncols, 128);
    SepArgs sa;
    sa.kc = kc;
    sa.Dc = Dc;
    sa.zbits = zbits + c0;
    sa.n_cont = dm.n;
    sa.mc = mc;
    sa.nb = nb;
    sa.col0 = c0;
    sa.colb = colb ? colb + c0 : nullptr;
    sa.ncols_dev = ncols_dev;
    FusedAcq fa;  // used by the kernel only when the factor has a single row tile (n + 1 <= 128)
    fa.aq = aq;
    fa.a = a + c0;
    fa.dmu = dmu + c0;
    fa.dvar = dvar + c0;
    fa.ncols = ncols;
    CUDA_TRY(launch_tma_lower_gen(m, sa, ncols_pad, need_grad ? panelV : nullptr, norm, mu, &fa, st));
    if (m.Mp1 > GEMM_BM)
      CUDA_TRY(launch_acq_epilogue(m, aq, norm, ncols_pad, mu, ncols, a + c0, dmu + c0, dvar + c0, nullptr, nullptr,
                                   ncols_dev, c0, st));
    if (need_grad) {
      CUDA_TRY(launch_tma_upper_grad(m, panelV, pl.C, sa, ncols, ncols_pad, dmu + c0, dvar + c0, S_part, st));
      CUDA_TRY(launch_sum_splits(m, S_part, m.Mp2 / GEMM_BM, &dm, ncols, ncols_pad, S_all + size_t(c0) * m.d_k, ncols_dev,
                                 c0, st));
    }
  }
  return PRB_OK;
}

uint64_t cont_mask(const DevSpace& sp) {
  uint64_t mask = 0;
  for (int i = 0; i < sp.n_cont; ++i) mask |= uint64_t(1) << sp.cont_cols[i];
  return mask;
}

// One fused step of the separable path on a single column panel (see prb_kernels.cu, "Fused optimisation step"):
// [clamp +] kc/Dc + sampler -> lower product (+ acquisition epilogue) -> upper product (+ gradient contraction) ->
// reductions [+ Adam] + EMA.  theta_in is read (and clamped to [lower, upper] if given); Xc receives the clamped rows.
int sep_step(const Model& m, const DevSpace& sp, const DevAcq& aq, const Plan& pl, void* ws, const BitMap& bm,
             const DimMap& dm, const double* theta_in, const double* lower, const double* upper, double* Xc, int b,
             int mc, const double* u_int, const double* u_cat, uint64_t seed, uint64_t offset, int estimator,
             double* ma_state, int update_ma, bool need_grad, const double* grad_out, double* out_value,
             double* out_grad, double* adam_theta, double* adam_m, double* adam_v, double lr, int* step_ptr,
             cudaStream_t st, double* a_out = nullptr) {
  const int ncols = b * mc;
  const int ncols_pad = round_up(ncols, 128);
  double* panelV = at<double>(ws, pl.off_panelV);
  double* norm = at<double>(ws, pl.off_norm);
  double* mu = at<double>(ws, pl.off_mu);
  double* S_part = at<double>(ws, pl.off_S_part);
  double* kc = at<double>(ws, pl.off_kc);
  double* Dc = at<double>(ws, pl.off_Dc);
  uint32_t* zbits = at<uint32_t>(ws, pl.off_zbits);
  uint8_t* z = at<uint8_t>(ws, pl.off_z);
  double* prob = at<double>(ws, pl.off_prob);
  double* a = a_out ? a_out : at<double>(ws, pl.off_a);
  double* dmu = at<double>(ws, pl.off_dmu);
  double* dvar = at<double>(ws, pl.off_dvar);
  CUDA_TRY(launch_sep_step_prep(m, sp, bm, theta_in, lower, upper, Xc, b, mc, u_int, u_cat, seed, offset, kc,
                                need_grad ? Dc : nullptr, need_grad ? z : nullptr, need_grad ? prob : nullptr, zbits, st));
  PdlScope pdl(!env_knobs().no_pdl);  // the remaining kernels of the step overlap their launch + prologue with their predecessor
  SepArgs sa;
  sa.kc = kc;
  sa.Dc = Dc;
  sa.zbits = zbits;
  sa.n_cont = dm.n;
  sa.mc = mc;
  sa.nb = b;
  sa.col0 = 0;
  sa.colb = nullptr;
  sa.ncols_dev = nullptr;
  FusedAcq fa;
  fa.aq = aq;
  fa.a = a;
  fa.dmu = dmu;
  fa.dvar = dvar;
  fa.ncols = ncols;
  CUDA_TRY(launch_tma_lower_gen(m, sa, ncols_pad, need_grad ? panelV : nullptr, norm, mu, &fa, st));
  if (m.Mp1 > GEMM_BM)
    CUDA_TRY(launch_acq_epilogue(m, aq, norm, ncols_pad, mu, ncols, a, dmu, dvar, nullptr, nullptr, nullptr, 0, st));
  if (need_grad) CUDA_TRY(launch_tma_upper_grad(m, panelV, pl.C, sa, ncols, ncols_pad, dmu, dvar, S_part, st));
  CUDA_TRY(launch_sep_step_finish(sp, dm, dm.n, 0, b, mc, a, z, prob, S_part, m.Mp2 / GEMM_BM, ncols_pad, estimator,
                                  ma_state,
                                  update_ma, grad_out, out_value, out_grad, need_grad ? 1 : 0, adam_theta, adam_m, adam_v,
                                  lr, step_ptr, m.done_ctr, st));
  return PRB_OK;
}

// The same fusion for kernel structures that do not separate (ordinals inside the ARD Matern, categorical kernels), on a
// single column panel: [clamp +] sampler -> k* panel -> lower product (+ acquisition epilogue when there is one row tile)
// -> upper product -> gradient contraction -> reductions [+ Adam] + EMA: six launches instead of twelve.
int generic_step(const Model& m, const DevSpace& sp, const DevAcq& aq, const Plan& pl, void* ws, const double* theta_in,
                 const double* lower, const double* upper, double* Xc, int b, int mc, const double* u_int,
                 const double* u_cat, uint64_t seed, uint64_t offset, int estimator, double* ma_state, int update_ma,
                 bool need_grad, const double* grad_out, double* out_value, double* out_grad, double* adam_theta,
                 double* adam_m, double* adam_v, double lr, int* step_ptr, cudaStream_t st, double* a_out = nullptr) {
  const int ncols = b * mc;
  const int ncols_pad = round_up(ncols, 128);
  double* panelA = at<double>(ws, pl.off_panelA);
  double* panelV = at<double>(ws, pl.off_panelV);
  double* norm = at<double>(ws, pl.off_norm);
  double* mu = at<double>(ws, pl.off_mu);
  double* S_part = at<double>(ws, pl.off_S_part);
  double* xk = at<double>(ws, pl.off_xk);
  uint8_t* z = at<uint8_t>(ws, pl.off_z);
  double* prob = at<double>(ws, pl.off_prob);
  double* a = a_out ? a_out : at<double>(ws, pl.off_a);
  double* dmu = at<double>(ws, pl.off_dmu);
  double* dvar = at<double>(ws, pl.off_dvar);
  const bool need_path = need_grad && sp.n_cont > 0;
  CUDA_TRY(launch_pr_sample(sp, theta_in, b, mc, u_int, u_cat, seed, offset, xk, nullptr, need_grad ? z : nullptr,
                            need_grad ? prob : nullptr, nullptr, nullptr, st, lower, upper, Xc));
  if (m.rff_m > 0) {
    // deterministic posterior sample (Thompson sampling): three launches -- clamp + sampler, feature evaluation (value and
    // gradient, written in the finish kernel's [dim][col] layout), finish
    if (aq.kind != PRB_ACQ_MEAN) return PRB_ERR_INVALID;
    CUDA_TRY(launch_rff_eval(m, xk, ncols, cont_mask(sp), need_path, a, S_part, nullptr, nullptr, st, ncols_pad));
    DimMap dmr;
    dmr.n = need_path ? sp.n_cont : 0;
    for (int i = 0; i < sp.n_cont; ++i) dmr.dims[i] = sp.cont_cols[i];
    CUDA_TRY(launch_sep_step_finish(sp, dmr, m.d_k, 1, b, mc, a, z, prob, S_part, 1, ncols_pad, estimator, ma_state,
                                    update_ma, grad_out, out_value, out_grad, need_grad ? 1 : 0, adam_theta, adam_m, adam_v,
                                    lr, step_ptr, m.done_ctr, st));
    return PRB_OK;
  }
  double* panelQ = at<double>(ws, pl.off_panelQ);
  GradGroup gq;
  const bool gq_ok = need_path && build_grad_group(m.spec, cont_mask(sp), &gq);
  PdlScope pdl(!env_knobs().no_pdl);  // the kernels below overlap their launch + prologue with their predecessor
  CUDA_TRY(launch_kstar_panel(m, xk, ncols, ncols_pad, panelA, st, gq_ok ? &gq : nullptr, panelQ));
  FusedAcq fa;
  fa.aq = aq;
  fa.a = a;
  fa.dmu = dmu;
  fa.dvar = dvar;
  fa.ncols = ncols;
  const bool fuse = m.Mp1 == GEMM_BM;
  CUDA_TRY(launch_tma_lower(m, panelA, pl.C, ncols_pad, need_path ? panelV : nullptr, norm, mu, fuse ? &fa : nullptr, st));
  if (!fuse)
    CUDA_TRY(launch_acq_epilogue(m, aq, norm, ncols_pad, mu, ncols, a, dmu, dvar, nullptr, nullptr, nullptr, 0, st));
  int jblocks = m.Mp2 / pl.grad_jrows;
  if (need_path && gq_ok) {
    CUDA_TRY(launch_tma_upper_gradq(m, panelV, pl.C, ncols, ncols_pad, gq, panelQ, xk, dmu, dvar, S_part, st));
    jblocks = m.Mp2 / GEMM_BM;
  } else if (need_path) {
    PdlScope plain(false);  // these two kernels do not carry the griddepcontrol protocol
    CUDA_TRY(launch_tma_upper(m, panelV, pl.C, ncols_pad, panelA, st));
    CUDA_TRY(launch_grad_contract(m, cont_mask(sp), xk, ncols, ncols_pad, panelA, dmu, dvar, S_part, pl.grad_jrows, st));
  }
  DimMap dm;
  dm.n = need_path ? sp.n_cont : 0;
  for (int i = 0; i < sp.n_cont; ++i) dm.dims[i] = sp.cont_cols[i];
  CUDA_TRY(launch_sep_step_finish(sp, dm, m.d_k, 1, b, mc, a, z, prob, S_part, jblocks, ncols_pad, estimator,
                                  ma_state, update_ma, grad_out, out_value, out_grad, need_grad ? 1 : 0, adam_theta,
                                  adam_m, adam_v, lr, step_ptr, m.done_ctr, st));
  return PRB_OK;
}

int mc_core(prb_model* model, const DevSpace& sp, const DevAcq& aq, const Plan& pl, void* ws, const double* theta,
            int b, int mc, const double* u_int, const double* u_cat, uint64_t seed, uint64_t offset, int estimator,
            double* ma_state, int update_ma, const double* grad_out, double* out_value, double* out_grad,
            double* out_acq_values, cudaStream_t st) {
  const Model& m = *reinterpret_cast<Model*>(model);
  const int64_t B = int64_t(b) * mc;
  double* xk = at<double>(ws, pl.off_xk);
  uint8_t* z = at<uint8_t>(ws, pl.off_z);
  double* prob = at<double>(ws, pl.off_prob);
  double* a = out_acq_values ? out_acq_values : at<double>(ws, pl.off_a);
  double* dmu = at<double>(ws, pl.off_dmu);
  double* dvar = at<double>(ws, pl.off_dvar);
  double* S_all = at<double>(ws, pl.off_S_all);
  const bool need_grad = grad_out != nullptr;
  const bool need_path = need_grad && sp.n_cont > 0;
  BitMap bm;
  DimMap dm;
  const bool sep = pl.off_kc != 0 && sep_space_ok(m, sp, &bm, &dm);
  int rc;
  const bool dedup = sep && m.opt_dedup && m.sep_nbits <= DEDUP_MAX_BITS;
  // experiment sizes: the whole call -- sampler, both products, acquisition, reductions, EMA -- is one kernel (prb_small.cu)
  int small_kind = 0;  // 1: separable kernel, 2: generic single-term kernel
  if (out_acq_values == nullptr && small_step_scratch_bytes(b, mc, sp.d) <= pl.small_bytes) {
    if (sep && !dedup && small_step_eligible(m, sp, dm, b, mc)) small_kind = 1;
    else if (!sep && small_gen_eligible(m, sp, &dm, b, mc)) small_kind = 2;
  }
  if (small_kind != 0) {
    // completion counters live behind the per-tile partial sums: zero the (small) scratch; launches re-arm them themselves
    CUDA_TRY(cudaMemsetAsync(at<char>(ws, pl.off_small), 0, small_step_scratch_bytes(b, mc, sp.d), st));
    CUDA_TRY(launch_small_step(m, small_kind == 2, sp, aq, bm, dm, theta, nullptr, nullptr, nullptr, b, mc, u_int, u_cat, seed, offset,
                               estimator, ma_state, (update_ma && ma_state) ? 1 : 0, need_grad, grad_out, out_value,
                               out_grad, nullptr, nullptr, nullptr, 0.0, nullptr, at<char>(ws, pl.off_small), st));
    return PRB_OK;
  }
  if (dedup) {
    // unique (restart, code) columns only; per-sample values are scattered back on request
    double* a_u = at<double>(ws, pl.off_a);
    CUDA_TRY(launch_pr_sample(sp, theta, b, mc, u_int, u_cat, seed, offset, nullptr, nullptr, nullptr,
                              need_grad ? prob : nullptr, &bm, at<uint32_t>(ws, pl.off_zbits), st));
    CUDA_TRY(launch_dedup(at<uint32_t>(ws, pl.off_zbits), b, mc, m.sep_nbits, at<int>(ws, pl.off_ucount),
                          at<int>(ws, pl.off_ustart), at<int>(ws, pl.off_utotal), at<uint32_t>(ws, pl.off_ucode),
                          at<int>(ws, pl.off_uweight), out_acq_values ? at<int>(ws, pl.off_inv) : nullptr,
                          at<uint32_t>(ws, pl.off_zbits_u), at<int>(ws, pl.off_colb), at<double>(ws, pl.off_wu), st));
    rc = run_engine_sep(m, aq, pl, ws, theta, sp.d, b, mc, dm, need_path, true, a_u, dmu, dvar, S_all, st);
    if (rc != PRB_OK) return rc;
    CUDA_TRY(launch_mc_reduce_dedup(sp, bm, b, mc, at<int>(ws, pl.off_ucount), at<int>(ws, pl.off_ustart),
                                    at<double>(ws, pl.off_wu), at<uint32_t>(ws, pl.off_zbits_u), a_u, prob, S_all,
                                    estimator, ma_state, grad_out, out_value, out_grad, st));
    if (out_acq_values)
      CUDA_TRY(launch_dedup_scatter(a_u, at<int>(ws, pl.off_ustart), at<int>(ws, pl.off_inv), b, mc, out_acq_values,
                                    st));
    if (update_ma && ma_state) CUDA_TRY(launch_ma_update(out_value, b, ma_state, st));
    return PRB_OK;
  }
  // One column panel (every experiment shape and the benchmark shape): the fused step -- 4-5 launches chained by
  // programmatic dependent launch (separable structure) or 6 (generic) instead of 8-12 stream-ordered ones.  The multi-
  // launch sequence below remains for multi-panel calls and as the cross-check path (PRB_OPT_UNFUSED).
  if (B <= pl.C && !m.opt_unfused && !env_knobs().no_fused_step) {
    const int upd = (update_ma && ma_state) ? 1 : 0;
    if (sep)
      return sep_step(m, sp, aq, pl, ws, bm, dm, theta, nullptr, nullptr, nullptr, b, mc, u_int, u_cat, seed, offset,
                      estimator, ma_state, upd, need_grad, grad_out, out_value, out_grad, nullptr, nullptr, nullptr, 0.0,
                      nullptr, st, out_acq_values);
    return generic_step(m, sp, aq, pl, ws, theta, nullptr, nullptr, nullptr, b, mc, u_int, u_cat, seed, offset, estimator,
                        ma_state, upd, need_grad, grad_out, out_value, out_grad, nullptr, nullptr, nullptr, 0.0, nullptr,
                        st, out_acq_values);
  }
  if (sep) {
    CUDA_TRY(launch_pr_sample(sp, theta, b, mc, u_int, u_cat, seed, offset, nullptr, nullptr, need_grad ? z : nullptr,
                              need_grad ? prob : nullptr, &bm, at<uint32_t>(ws, pl.off_zbits), st));
    rc = run_engine_sep(m, aq, pl, ws, theta, sp.d, b, mc, dm, need_path, false, a, dmu, dvar, S_all, st);
  } else {
    CUDA_TRY(launch_pr_sample(sp, theta, b, mc, u_int, u_cat, seed, offset, xk, nullptr, need_grad ? z : nullptr,
                              need_grad ? prob : nullptr, nullptr, nullptr, st));
    rc = run_engine(m, aq, pl, ws, xk, B, need_path, cont_mask(sp), a, dmu, dvar, S_all, nullptr, nullptr, st);
  }
  if (rc != PRB_OK) return rc;
  CUDA_TRY(launch_mc_reduce(sp, b, mc, a, z, prob, S_all, estimator, ma_state, grad_out, out_value, out_grad, st));
  if (update_ma && ma_state) CUDA_TRY(launch_ma_update(out_value, b, ma_state, st));
  return PRB_OK;
}

// Run `one_iter` maxiter times.  With a real stream the step is captured once into a CUDA graph and replayed (removes the
// per-launch host latency: the reference spends its time exactly there at experiment sizes, 800 host-driven steps per BO
// iteration), else it is issued as plain launches.  The executable graph is cached on the model
// (`slot`): a later call re-captures its step -- a host-side recording of a handful of launches -- and updates the cached
// executable in place (cudaGraphExecUpdate only affects FUTURE launches, so replays still in flight are untouched).  No
// instantiation and no stream synchronisation in steady state; an executable whose topology no longer matches is parked
// on the model and released by prb_model_destroy.
template <typename Iter>
int replay_steps(Model& m, int slot, Iter one_iter, int maxiter, cudaStream_t st) {
  bool graphed = false;
  if (st != nullptr && maxiter >= 4) {
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      const unsigned long long l0 = g_launches.load(std::memory_order_relaxed);
      int r = one_iter();
      const unsigned long long per_iter = g_launches.load(std::memory_order_relaxed) - l0;
      cudaError_t e = cudaStreamEndCapture(st, &graph);
      if (r == PRB_OK && e == cudaSuccess && graph != nullptr) {
        cudaGraphExec_t& exec = m.graph_exec[slot];
        bool ready = false;
        if (exec != nullptr) {
          cudaGraphExecUpdateResultInfo info;
          memset(&info, 0, sizeof(info));
          if (cudaGraphExecUpdate(exec, graph, &info) == cudaSuccess) {
            ready = true;
          } else {
            (void)cudaGetLastError();
            if (m.n_graph_dead < int(sizeof(m.graph_dead) / sizeof(m.graph_dead[0]))) {
              m.graph_dead[m.n_graph_dead++] = exec;
            } else {  // parking lot full (a caller alternating between many topologies): drain, then release
              cudaStreamSynchronize(st);
              cudaGraphExecDestroy(exec);
            }
            exec = nullptr;
          }
        }
        if (!ready) ready = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
        if (ready) {
          cudaError_t le = cudaSuccess;
          for (int it = 0; it < maxiter && le == cudaSuccess; ++it) le = cudaGraphLaunch(exec, st);
          if (le != cudaSuccess) {
            cudaGraphDestroy(graph);
            return cuda_fail(le, "cudaGraphLaunch");
          }
          graphed = true;
          count_launch(int(per_iter) * (maxiter - 1));  // replays of the captured step
        } else {
          exec = nullptr;
        }
      }
      if (graph) cudaGraphDestroy(graph);
      (void)cudaGetLastError();
    }
  }
  if (!graphed) {
    for (int it = 0; it < maxiter; ++it) {
      int rc = one_iter();
      if (rc != PRB_OK) return rc;
    }
  }
  return PRB_OK;
}

template <typename Core>
int adam_graph_loop(Model& m, int slot, Core core, double* theta, const double* lower, const double* upper, int b, int d, double lr, int maxiter,
                    int step0, double* adam_state, double* out_value, double* Xc, double* grad, double* ones, double* val,
                    int* step, cudaStream_t st) {
  const int total = b * d;
  CUDA_TRY(launch_fill(ones, 1.0, b, st));
  CUDA_TRY(cudaMemcpyAsync(step, &step0, sizeof(int), cudaMemcpyHostToDevice, st));
  auto one_iter = [&]() -> int {
    CUDA_TRY(launch_clamp(theta, lower, upper, b, d, Xc, st));
    int r = core(Xc, ones, val, grad);
    if (r != PRB_OK) return r;
    CUDA_TRY(launch_adam_step(theta, grad, adam_state, adam_state + total, total, lr, step, st));
    return PRB_OK;
  };
  int rc = replay_steps(m, slot, one_iter, maxiter, st);
  if (rc != PRB_OK) return rc;
  CUDA_TRY(launch_clamp(theta, lower, upper, b, d, theta, st));
  return core(theta, nullptr, out_value, nullptr);
}

// The library carries sm_100a code only.  (Attribute queries: cudaGetDeviceProperties costs milliseconds per call.)
int check_device(int dev) {
  int major = 0, minor = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "device %d is sm_%d%d; libprb200 is built for sm_100a only", dev, major, minor);
    return PRB_ERR_NO_DEVICE;
  }
  return PRB_OK;
}

}  // namespace

// =====================================================================================================================
extern "C" {

int prb_version(void) { return PRB_VERSION; }

uint64_t prb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int prb_abi_sizeof(int which) {
  switch (which) {
    case 0: return int(sizeof(prb_factor));
    case 1: return int(sizeof(prb_term));
    case 2: return int(sizeof(prb_model_desc));
    case 3: return int(sizeof(prb_space_desc));
    case 4: return int(sizeof(prb_acq_desc));
    case 5: return int(sizeof(prb_multi_acq_desc));
    default: return PRB_ERR_INVALID;
  }
}

int prb_debug_small_timestamps(long long* out16) {
  if (!out16) return PRB_ERR_INVALID;
  CUDA_TRY(small_step_timestamps(out16));
  return PRB_OK;
}

int prb_debug_fp64_peak(int mode, double* out_tflops, void* stream) {
  if (!out_tflops || mode < 0 || mode > 5) return PRB_ERR_INVALID;
  int dev = -1, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return PRB_ERR_NO_DEVICE;
  if (int rc_dev = check_device(dev)) return rc_dev;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* scratch = nullptr;
  CUDA_TRY(cudaMalloc(&scratch, size_t(2) * sms * 256 * sizeof(double)));
  cudaError_t e = run_fp64_peak(mode, 1 << 15, scratch, sms, out_tflops, static_cast<cudaStream_t>(stream));
  cudaFree(scratch);
  if (e != cudaSuccess) return cuda_fail(e, "prb_debug_fp64_peak");
  return PRB_OK;
}

int prb_profile_enable(int on) {
  g_prof_on = on != 0;
  return PRB_OK;
}

int prb_profile_read(double* ms_by_kernel, uint64_t* launches_by_kernel, int32_t n_slots) {
  if (n_slots < PRB_PROFILE_SLOTS) return PRB_ERR_INVALID;
  for (int i = 0; i < n_slots; ++i) {
    if (ms_by_kernel) ms_by_kernel[i] = 0.0;
    if (launches_by_kernel) launches_by_kernel[i] = 0;
  }
  cudaError_t first = cudaSuccess;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (const auto& r : g_prof) {
    cudaError_t e = cudaEventSynchronize(r.end);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.beg, r.end);
    if (e == cudaSuccess) {
      if (ms_by_kernel) ms_by_kernel[r.id] += double(ms);
      if (launches_by_kernel) launches_by_kernel[r.id] += 1;
    } else if (first == cudaSuccess) {
      first = e;
    }
    cudaEventDestroy(r.beg);
    cudaEventDestroy(r.end);
  }
  g_prof.clear();
  if (first != cudaSuccess) return cuda_fail(first, "prb_profile_read");
  return PRB_OK;
}

const char* prb_status_string(int status) {
  switch (status) {
    case PRB_OK: return "ok";
    case PRB_ERR_INVALID: return "invalid argument";
    case PRB_ERR_UNSUPPORTED: return "exceeds a compile-time limit of libprb200";
    case PRB_ERR_WORKSPACE: return "workspace too small";
    case PRB_ERR_CUDA: return "CUDA error";
    case PRB_ERR_NO_DEVICE: return "no usable sm_100 device";
    case PRB_ERR_NOT_PD: return "matrix is not positive definite";
    default: return "unknown status";
  }
}

const char* prb_last_cuda_error(void) { return g_cuda_err; }

static int model_create_impl(prb_model** out, const prb_model_desc* desc, const double* y, const double* noise,
                             double* out_Linv, double* out_alpha, void* stream);

int prb_model_create(prb_model** out, const prb_model_desc* desc, void* stream) {
  if (!desc || !desc->Linv || !desc->alpha) return PRB_ERR_INVALID;
  return model_create_impl(out, desc, nullptr, nullptr, nullptr, nullptr, stream);
}

int prb_model_create_from_data(prb_model** out, const prb_model_desc* desc, const double* y, const double* noise,
                               double* out_Linv, double* out_alpha, void* stream) {
  if (!y || !noise) return PRB_ERR_INVALID;
  return model_create_impl(out, desc, y, noise, out_Linv, out_alpha, stream);
}

static int model_create_impl(prb_model** out, const prb_model_desc* desc, const double* y, const double* noise,
                             double* out_Linv, double* out_alpha, void* stream) {
  const bool from_data = y != nullptr;
  if (!out || !desc || desc->n < 1 || desc->d_k < 1 || !desc->X_train) return PRB_ERR_INVALID;
  if (desc->d_k > PRB_MAX_DIMS) return PRB_ERR_UNSUPPORTED;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return PRB_ERR_NO_DEVICE;
  if (int rc_dev = check_device(dev)) return rc_dev;
  Model* m = new (std::nothrow) Model();
  if (!m) return PRB_ERR_INVALID;
  memset(m, 0, sizeof(Model));
  int rc = make_spec(desc, &m->spec, &m->kss);
  if (rc != PRB_OK) {
    delete m;
    return rc;
  }
  m->n = desc->n;
  m->d_k = desc->d_k;
  m->Kp = round_up(desc->n, GEMM_BK);
  m->Mp1 = round_up(desc->n + 1, GEMM_BM);
  m->Mp2 = round_up(desc->n, GEMM_BM);
  m->mean_const = desc->mean_const;
  m->device = dev;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  // Stream-ordered pool allocations throughout: four cudaMalloc calls for the model arrays cost 630 us per build at
  // n = 1000 -- a third of the whole build -- and cudaMalloc / cudaFree of the factorisation scratch more than the
  // factorisation itself.  The pool keeps what it has handed out (release threshold raised once), so the next BO
  // iteration's build reuses it.  All device arrays of the model live in ONE block.
  static std::once_flag pool_once;
  std::call_once(pool_once, [dev] {
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = uint64_t(2) << 30;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    (void)cudaGetLastError();
  });
  const int n_pad = m->Mp1 > m->Mp2 ? m->Mp1 : m->Mp2;
  size_t off = 0;
  auto carve = [&off](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~size_t(255);
    return o;
  };
  const size_t o_A1 = carve(size_t(m->Mp1) * m->Kp * sizeof(double));
  const size_t o_A2 = carve(size_t(m->Mp2) * m->Kp * sizeof(double));
  const size_t o_X = carve(size_t(m->n) * m->d_k * sizeof(double));
  const size_t o_alpha = carve(size_t(m->n) * sizeof(double));
  const size_t o_apad = carve(size_t(n_pad) * sizeof(double));
  const size_t o_Zb = carve(size_t(n_pad) * sizeof(uint32_t));
  const size_t o_tab = carve(64 * sizeof(double));
  const size_t o_ctr = carve(64);
  const size_t o_flags = carve(64);  // [0] Cholesky info, [1] separable-structure check
  const size_t o_ones = carve(MODEL_ONES * sizeof(double));
  if ((e = cudaMallocAsync(&m->block, off, st)) != cudaSuccess) {
    delete m;
    return cuda_fail(e, "cudaMallocAsync(model)");
  }
  char* blk = static_cast<char*>(m->block);
  m->A1 = reinterpret_cast<double*>(blk + o_A1);
  m->A2 = reinterpret_cast<double*>(blk + o_A2);
  m->Xtr = reinterpret_cast<double*>(blk + o_X);
  m->alpha = reinterpret_cast<double*>(blk + o_alpha);
  m->alpha_pad = reinterpret_cast<double*>(blk + o_apad);
  m->Zbits = reinterpret_cast<uint32_t*>(blk + o_Zb);
  m->sep_tab = reinterpret_cast<double*>(blk + o_tab);
  m->done_ctr = reinterpret_cast<unsigned int*>(blk + o_ctr);
  int* flags_dev = reinterpret_cast<int*>(blk + o_flags);
  m->ones = reinterpret_cast<double*>(blk + o_ones);
  e = cudaMemcpyAsync(m->Xtr, desc->X_train, size_t(m->n) * m->d_k * sizeof(double), cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(blk + o_ctr, 0, 128, st);  // completion counter + both flags
  if (e == cudaSuccess) e = cudaMemsetAsync(m->Zbits, 0, size_t(n_pad) * sizeof(uint32_t), st);
  if (e == cudaSuccess) e = launch_fill(m->ones, 1.0, MODEL_ONES, st);
  if (e == cudaSuccess) e = tma_init();
  if (e == cudaSuccess) e = small_step_init();
  // separable candidates: one product term of two Matern-5/2 factors, one of them isotropic over <= 32 dims.  The check that
  // the factor's training columns really are {0, 1} runs on the device; its flag is read back with the factorisation's.
  int sep_cand[2] = {0, 0};
  if (m->spec.n_terms == 1 && !m->spec.t[0].additive && m->spec.t[0].n_factors == 2) {
    const DevTerm& T = m->spec.t[0];
    for (int fb = 0; fb < 2; ++fb) {
      const DevFactor& F = T.f[fb];
      const DevFactor& O = T.f[1 - fb];
      bool cand = F.kind == PRB_FACTOR_MATERN52 && O.kind == PRB_FACTOR_MATERN52 && F.power == 1 && O.power == 1 &&
                  F.n_dims <= 32;
      for (int i = 1; i < F.n_dims && cand; ++i) cand = F.inv_ls[i] == F.inv_ls[0];
      for (int i = 0; i < F.n_dims && cand; ++i)
        for (int k = 0; k < O.n_dims && cand; ++k) cand = F.dims[i] != O.dims[k];
      sep_cand[fb] = cand ? 1 : 0;
    }
  }
  auto start_sep_check = [&](int fb) -> cudaError_t {
    m->sep_bin_factor = fb;
    m->sep_cont_factor = 1 - fb;
    m->sep_nbits = m->spec.t[0].f[fb].n_dims;
    cudaError_t r = cudaMemsetAsync(flags_dev + 1, 0, sizeof(int), st);
    if (r == cudaSuccess) r = cudaMemsetAsync(m->Zbits, 0, size_t(n_pad) * sizeof(uint32_t), st);
    if (r == cudaSuccess) r = launch_sep_model_prep(*m, n_pad, flags_dev + 1, st);
    return r;
  };
  m->sep_ok = 0;
  const int first = sep_cand[0] ? 0 : (sep_cand[1] ? 1 : -1);
  if (e == cudaSuccess && first >= 0) e = start_sep_check(first);
  int flags_host[2] = {0, 1};  // Cholesky info (0 = fine), separable check (0 = columns are {0, 1})
  if (e == cudaSuccess && from_data) {
    // K(X, X) by the library's own kernel, then the factorisation, the inverse factor, alpha and the packing (prb_chol.cu)
    double* Kmat = nullptr;
    void* scratch = nullptr;
    if ((e = cudaMallocAsync(&Kmat, size_t(m->n) * m->n * sizeof(double), st)) == cudaSuccess &&
        (e = cudaMallocAsync(&scratch, chol_scratch_bytes(m->n), st)) == cudaSuccess &&
        (e = launch_kernel_matrix(m->spec, m->Xtr, m->n, m->Xtr, m->n, Kmat, st)) == cudaSuccess)
      e = build_gp_state(Kmat, noise, y, m->mean_const, m->n, m->Kp, m->Mp1, m->Mp2, m->A1, m->A2, m->alpha, out_Linv,
                         scratch, flags_dev, st);
    if (e == cudaSuccess && out_alpha)
      e = cudaMemcpyAsync(out_alpha, m->alpha, size_t(m->n) * sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (Kmat) cudaFreeAsync(Kmat, st);
    if (scratch) cudaFreeAsync(scratch, st);
  } else if (e == cudaSuccess) {
    e = cudaMemcpyAsync(m->alpha, desc->alpha, size_t(m->n) * sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) {
      dim3 grid((m->Kp + 127) / 128, m->Mp1 > m->Mp2 ? m->Mp1 : m->Mp2);
      pack_state_kernel<<<grid, 128, 0, st>>>(desc->Linv, desc->alpha, m->n, m->Kp, m->Mp1, m->Mp2, m->A1, m->A2);
      e = cudaGetLastError();
      count_launch();
    }
  }
  if (e == cudaSuccess) e = launch_pad_copy(m->alpha, m->n, n_pad, m->alpha_pad, st);
  // ONE synchronisation: the factorisation's status and the separable-structure check come back together
  if (e == cudaSuccess) e = cudaMemcpyAsync(flags_host, flags_dev, 2 * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess && from_data && flags_host[0] != 0) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "K + diag(noise) is not positive definite (pivot %d)", flags_host[0] - 1);
    prb_model_destroy(reinterpret_cast<prb_model*>(m));
    return PRB_ERR_NOT_PD;
  }
  if (e == cudaSuccess && first >= 0) {
    if (flags_host[1] == 0) {
      m->sep_ok = 1;  // the factor's training columns are exactly {0, 1}
    } else if (first == 0 && sep_cand[1]) {  // rare: the other factor order is a structural candidate too
      int bad = 1;
      if ((e = start_sep_check(1)) == cudaSuccess &&
          (e = cudaMemcpyAsync(&bad, flags_dev + 1, sizeof(int), cudaMemcpyDeviceToHost, st)) == cudaSuccess &&
          (e = cudaStreamSynchronize(st)) == cudaSuccess && !bad)
        m->sep_ok = 1;
    }
  }
  // second-generation engine: TMA descriptors of the packed factors
  if (e == cudaSuccess) e = make_operand_map(&m->mapA1, m->A1, m->Mp1, m->Kp, m->Kp);
  if (e == cudaSuccess) e = make_operand_map(&m->mapA2, m->A2, m->Mp2, m->Kp, m->Kp);
  if (e == cudaSuccess) e = make_operand_map3(&m->mapA1k, m->A1, m->Mp1, m->Kp, m->Kp);
  if (e == cudaSuccess) e = make_operand_map3(&m->mapA2k, m->A2, m->Mp2, m->Kp, m->Kp);
  if (e != cudaSuccess) {
    prb_model_destroy(reinterpret_cast<prb_model*>(m));
    return cuda_fail(e, "prb_model_create");
  }
  *out = reinterpret_cast<prb_model*>(m);
  return PRB_OK;
}

int prb_rff_model_create(prb_model** out, int32_t d_k, int32_t n_features, const double* W, const double* bias,
                         const double* weights, double scale, void* stream) {
  if (!out || d_k < 1 || n_features < 1 || !W || !bias || !weights) return PRB_ERR_INVALID;
  if (d_k > PRB_MAX_DIMS) return PRB_ERR_UNSUPPORTED;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return PRB_ERR_NO_DEVICE;
  if (int rc_dev = check_device(dev)) return rc_dev;
  Model* m = new (std::nothrow) Model();
  if (!m) return PRB_ERR_INVALID;
  memset(m, 0, sizeof(Model));
  // the plan arithmetic (workspace sizes) sees a one-point model; none of the GP buffers exists
  m->n = 1;
  m->d_k = d_k;
  m->Kp = round_up(1, GEMM_BK);
  m->Mp1 = GEMM_BM;
  m->Mp2 = GEMM_BM;
  m->device = dev;
  m->spec.d_k = d_k;
  m->rff_m = n_features;
  m->rff_scale = scale;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t fm = size_t(n_features) * sizeof(double);
  cudaError_t e;
  if ((e = cudaMalloc(&m->rff_W, fm * d_k)) == cudaSuccess && (e = cudaMalloc(&m->rff_b, fm)) == cudaSuccess &&
      (e = cudaMalloc(&m->rff_w, fm)) == cudaSuccess && (e = cudaMalloc(&m->done_ctr, 64)) == cudaSuccess &&
      (e = cudaMemsetAsync(m->done_ctr, 0, 64, st)) == cudaSuccess &&
      (e = cudaMemcpyAsync(m->rff_W, W, fm * d_k, cudaMemcpyDeviceToDevice, st)) == cudaSuccess &&
      (e = cudaMemcpyAsync(m->rff_b, bias, fm, cudaMemcpyDeviceToDevice, st)) == cudaSuccess)
    e = cudaMemcpyAsync(m->rff_w, weights, fm, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) {
    prb_model_destroy(reinterpret_cast<prb_model*>(m));
    return cuda_fail(e, "prb_rff_model_create");
  }
  *out = reinterpret_cast<prb_model*>(m);
  return PRB_OK;
}

int prb_model_destroy(prb_model* model) {
  if (!model) return PRB_OK;
  Model* m = reinterpret_cast<Model*>(model);
  // synchronise the device first: no replay of a cached graph is in flight past this point
  if (m->block)
    cudaDeviceSynchronize();
  else
    cudaFree(m->rff_W);
  for (int i = 0; i < 2; ++i)
    if (m->graph_exec[i]) cudaGraphExecDestroy(m->graph_exec[i]);
  for (int i = 0; i < m->n_graph_dead; ++i) cudaGraphExecDestroy(m->graph_dead[i]);
  if (m->block) {
    cudaFree(m->block);  // pool memory: synchronises like any cudaFree, then goes back to the pool for the next build
  } else {
    cudaFree(m->done_ctr);  // RFF sample models own their arrays one by one
    cudaFree(m->rff_b);
    cudaFree(m->rff_w);
  }
  delete m;
  return PRB_OK;
}

int prb_model_set_option(prb_model* model, int32_t option, int64_t value) {
  if (!model) return PRB_ERR_INVALID;
  Model* m = reinterpret_cast<Model*>(model);
  switch (option) {
    case PRB_OPT_DEDUP:
      m->opt_dedup = value != 0;
      return PRB_OK;
    case PRB_OPT_UNFUSED:
      m->opt_unfused = value != 0;
      return PRB_OK;
    default:
      return PRB_ERR_INVALID;
  }
}

int prb_model_n(const prb_model* model) { return model ? reinterpret_cast<const Model*>(model)->n : PRB_ERR_INVALID; }
int prb_model_dk(const prb_model* model) {
  return model ? reinterpret_cast<const Model*>(model)->d_k : PRB_ERR_INVALID;
}

size_t prb_workspace_bytes(const prb_model* model, int64_t n_points, int32_t n_restarts, int32_t d_theta,
                           int32_t chunk_cols) {
  if (!model || n_points < 0 || d_theta < 1 || n_restarts < 0) return 0;
  return make_plan(*reinterpret_cast<const Model*>(model), n_points, n_restarts, d_theta, chunk_cols).total;
}

int prb_philox_uniform(uint64_t seed, uint64_t offset, int32_t mc, int32_t n_cols, double* out, void* stream) {
  if (mc < 0 || n_cols < 0 || (!out && mc * n_cols > 0)) return PRB_ERR_INVALID;
  CUDA_TRY(launch_philox_fill(seed, offset, mc, n_cols, out, static_cast<cudaStream_t>(stream)));
  return PRB_OK;
}

int prb_pr_sample(const prb_space_desc* space, const double* theta, int32_t b, int32_t mc, const double* u_int,
                  const double* u_cat, uint64_t seed, uint64_t offset, double* out_tilde_x, uint8_t* out_z,
                  double* out_prob, void* stream) {
  DevSpace sp;
  int rc = make_space(space, &sp);
  if (rc != PRB_OK) return rc;
  if (!theta || b < 0 || mc < 0) return PRB_ERR_INVALID;
  CUDA_TRY(launch_pr_sample(sp, theta, b, mc, u_int, u_cat, seed, offset, nullptr, out_tilde_x, out_z, out_prob,
                            nullptr, nullptr, static_cast<cudaStream_t>(stream)));
  return PRB_OK;
}

int prb_sample_candidates(const prb_space_desc* space, const double* theta, int32_t b, const double* u, uint64_t seed,
                          uint64_t offset, double* out_x, double* out_xk, void* stream) {
  DevSpace sp;
  int rc = make_space(space, &sp);
  if (rc != PRB_OK) return rc;
  if (!theta || b < 0 || (!out_x && !out_xk)) return PRB_ERR_INVALID;
  CUDA_TRY(launch_sample_candidates(sp, theta, b, u, seed, offset, out_x, out_xk, static_cast<cudaStream_t>(stream)));
  return PRB_OK;
}

int prb_finalize_candidates(prb_model* model, const prb_space_desc* space, const prb_acq_desc* acq, const double* theta,
                            int32_t b, const double* u, uint64_t seed, uint64_t offset, double* out_x, double* out_acq,
                            double* out_best, int32_t* out_best_index, void* workspace, size_t workspace_bytes,
                            void* stream) {
  if (!model || !theta || !out_x || b < 1 || !workspace) return PRB_ERR_INVALID;
  const Model& m = *reinterpret_cast<Model*>(model);
  DevSpace sp;
  DevAcq aq;
  int rc = make_space(space, &sp);
  if (rc != PRB_OK) return rc;
  if ((rc = make_acq(acq, &aq)) != PRB_OK) return rc;
  if (sp.d_k != m.d_k) return PRB_ERR_INVALID;
  const Plan pl = make_plan(m, b, 0, m.d_k, 0);
  if (workspace_bytes < pl.total) return PRB_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* xk = at<double>(workspace, pl.off_xk);
  double* a = out_acq ? out_acq : at<double>(workspace, pl.off_a);
  CUDA_TRY(launch_sample_candidates(sp, theta, b, u, seed, offset, out_x, xk, st));
  rc = run_engine(m, aq, pl, workspace, xk, b, false, uint64_t(0), a, at<double>(workspace, pl.off_dmu),
                  at<double>(workspace, pl.off_dvar), nullptr, nullptr, nullptr, st);
  if (rc != PRB_OK) return rc;
  if (out_best || out_best_index) CUDA_TRY(launch_best_row(a, out_x, b, sp.d, out_best, out_best_index, st));
  return PRB_OK;
}

int prb_trust_region_bounds(const double* X, const double* Y, int32_t n, int32_t d, int32_t n_constraints, double length,
                            const double* lower, const double* upper, const uint8_t* reset01, double* out_bounds,
                            int32_t* out_center_index, void* stream) {
  if (!X || !Y || n < 1 || d < 1 || n_constraints < 0 || !lower || !upper || !out_bounds || !(length > 0.0))
    return PRB_ERR_INVALID;
  CUDA_TRY(launch_tr_bounds(X, Y, n, d, n_constraints, length, lower, upper, reset01, out_bounds, out_center_index,
                            static_cast<cudaStream_t>(stream)));
  return PRB_OK;
}

int prb_analytic_enumerate(const prb_space_desc* space, const double* theta, int32_t b, const int32_t* options,
                           int32_t n_options, double* out_x_all, double* out_probs, void* stream) {
  DevSpace sp;
  int rc = make_space(space, &sp);
  if (rc != PRB_OK) return rc;
  if (!theta || !options || b < 0 || n_options < 1) return PRB_ERR_INVALID;
  CUDA_TRY(launch_analytic_build(sp, theta, b, n_options, options, out_x_all, out_probs,
                                 static_cast<cudaStream_t>(stream)));
  return PRB_OK;
}

int prb_kernel_matrix(const prb_model_desc* desc, const double* xk, int64_t n_points, double* out, void* stream) {
  if (!desc || !xk || !out || desc->n < 1 || desc->d_k < 1 || !desc->X_train || n_points < 0) return PRB_ERR_INVALID;
  if (desc->d_k > PRB_MAX_DIMS) return PRB_ERR_UNSUPPORTED;
  DevSpec spec;
  double kss;
  int rc = make_spec(desc, &spec, &kss);
  if (rc != PRB_OK) return rc;
  CUDA_TRY(launch_kernel_matrix(spec, desc->X_train, desc->n, xk, int(n_points), out,
                                static_cast<cudaStream_t>(stream)));
  return PRB_OK;
}

int prb_acq_points(prb_model* model, const prb_acq_desc* acq, const double* xk, int64_t n_points, double* out_acq,
                   double* out_mean, double* out_var, double* out_dacq_dx, void* workspace, size_t workspace_bytes,
                   int32_t chunk_cols, void* stream) {
  if (!model || !xk || n_points < 0 || !workspace) return PRB_ERR_INVALID;
  const Model& m = *reinterpret_cast<Model*>(model);
  DevAcq aq;
  int rc = make_acq(acq, &aq);
  if (rc != PRB_OK) return rc;
  if (n_points == 0) return PRB_OK;
  const Plan pl = make_plan(m, n_points, 0, m.d_k, chunk_cols);
  if (workspace_bytes < pl.total) return PRB_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* a = out_acq ? out_acq : at<double>(workspace, pl.off_a);
  uint64_t mask = 0;
  for (int t = 0; t < m.spec.n_terms; ++t)
    for (int f = 0; f < m.spec.t[t].n_factors; ++f)
      if (m.spec.t[t].f[f].kind == PRB_FACTOR_MATERN52)
        for (int i = 0; i < m.spec.t[t].f[f].n_dims; ++i) mask |= uint64_t(1) << m.spec.t[t].f[f].dims[i];
  if (m.rff_m > 0) mask = m.d_k >= 64 ? ~uint64_t(0) : ((uint64_t(1) << m.d_k) - 1u);  // a feature map is smooth in every input
  return run_engine(m, aq, pl, workspace, xk, n_points, out_dacq_dx != nullptr, mask, a,
                    at<double>(workspace, pl.off_dmu), at<double>(workspace, pl.off_dvar), out_dacq_dx, out_mean,
                    out_var, st);
}

int prb_mc_value_grad(prb_model* model, const prb_space_desc* space, const prb_acq_desc* acq, const double* theta,
                      int32_t b, int32_t mc, const double* u_int, const double* u_cat, uint64_t seed,
                      uint64_t offset, int32_t estimator, double* ma_state, int32_t update_ma,
                      const double* grad_out, double* out_value, double* out_grad, double* out_acq_values,
                      void* workspace, size_t workspace_bytes, int32_t chunk_cols, void* stream) {
  if (!model || !theta || !out_value || b < 0 || mc < 1 || !workspace) return PRB_ERR_INVALID;
  if (grad_out && !out_grad) return PRB_ERR_INVALID;
  if (estimator != PRB_EST_REINFORCE && estimator != PRB_EST_REINFORCE_MA) return PRB_ERR_INVALID;
  if ((estimator == PRB_EST_REINFORCE_MA || update_ma) && !ma_state) return PRB_ERR_INVALID;
  const Model& m = *reinterpret_cast<Model*>(model);
  DevSpace sp;
  DevAcq aq;
  int rc = make_space(space, &sp);
  if (rc != PRB_OK) return rc;
  if ((rc = make_acq(acq, &aq)) != PRB_OK) return rc;
  if (sp.d_k != m.d_k) return PRB_ERR_INVALID;
  if (b == 0) return PRB_OK;
  const Plan pl = make_plan(m, int64_t(b) * mc, b, sp.d, chunk_cols);
  if (workspace_bytes < pl.total) return PRB_ERR_WORKSPACE;
  return mc_core(model, sp, aq, pl, workspace, theta, b, mc, u_int, u_cat, seed, offset, estimator, ma_state,
                 update_ma, grad_out, out_value, out_grad, out_acq_values, static_cast<cudaStream_t>(stream));
}

int prb_mc_value_grad_host(prb_model* model, const prb_space_desc* space, const prb_acq_desc* acq,
                           const double* theta_host, int32_t b, int32_t mc, const double* u_int, const double* u_cat,
                           uint64_t seed, uint64_t offset, int32_t estimator, double* ma_state, int32_t update_ma,
                           int32_t want_grad, double* out_value_host, double* out_grad_host, void* workspace,
                           size_t workspace_bytes, int32_t synchronize, void* stream) {
  if (!model || !theta_host || !out_value_host || b < 1 || mc < 1 || !workspace) return PRB_ERR_INVALID;
  if (want_grad && !out_grad_host) return PRB_ERR_INVALID;
  if (estimator != PRB_EST_REINFORCE && estimator != PRB_EST_REINFORCE_MA) return PRB_ERR_INVALID;
  if ((estimator == PRB_EST_REINFORCE_MA || update_ma) && !ma_state) return PRB_ERR_INVALID;
  const Model& m = *reinterpret_cast<Model*>(model);
  DevSpace sp;
  DevAcq aq;
  int rc = make_space(space, &sp);
  if (rc != PRB_OK) return rc;
  if ((rc = make_acq(acq, &aq)) != PRB_OK) return rc;
  if (sp.d_k != m.d_k) return PRB_ERR_INVALID;
  const Plan pl = make_plan(m, int64_t(b) * mc, b, sp.d, 0);
  if (workspace_bytes < pl.total) return PRB_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // device staging inside the caller's workspace: theta, value, gradient, unit upstream gradient
  double* theta_dev = at<double>(workspace, pl.off_Xc);
  double* grad_dev = at<double>(workspace, pl.off_grad);
  // value staged right behind the gradient rows: a caller whose host buffers are laid out the same way ([b, d] gradient, then
  // [b] values, e.g. two views of one pinned block) gets both back with ONE device->host copy
  const bool fits = size_t(b) * (sp.d + 1) <= size_t(pl.B_pad) * sp.d;  // the gradient region holds B_pad x d doubles
  double* value_dev = (want_grad && fits) ? grad_dev + size_t(b) * sp.d : at<double>(workspace, pl.off_val);
  const bool resident_ones = m.ones != nullptr && b <= MODEL_ONES;
  double* ones = resident_ones ? m.ones : at<double>(workspace, pl.off_ones);
  const size_t row_bytes = size_t(b) * sp.d * sizeof(double);
  const bool one_copy = want_grad && fits && out_value_host == out_grad_host + size_t(b) * sp.d;
  auto sequence = [&](cudaStream_t s) -> int {
    CUDA_TRY(cudaMemcpyAsync(theta_dev, theta_host, row_bytes, cudaMemcpyHostToDevice, s));
    if (want_grad && !resident_ones) CUDA_TRY(launch_fill(ones, 1.0, b, s));
    int r = mc_core(model, sp, aq, pl, workspace, theta_dev, b, mc, u_int, u_cat, seed, offset, estimator, ma_state, update_ma,
                    want_grad ? ones : nullptr, value_dev, want_grad ? grad_dev : nullptr, nullptr, s);
    if (r != PRB_OK) return r;
    if (one_copy) {
      CUDA_TRY(cudaMemcpyAsync(out_grad_host, grad_dev, row_bytes + size_t(b) * sizeof(double), cudaMemcpyDeviceToHost, s));
    } else {
      CUDA_TRY(cudaMemcpyAsync(out_value_host, value_dev, size_t(b) * sizeof(double), cudaMemcpyDeviceToHost, s));
      if (want_grad) CUDA_TRY(cudaMemcpyAsync(out_grad_host, grad_dev, row_bytes, cudaMemcpyDeviceToHost, s));
    }
    return PRB_OK;
  };
  // (A cached executable graph of this whole sequence -- one cudaGraphLaunch instead of ~8 runtime calls -- was measured and
  // changes nothing: 0.645 vs 0.648 ms per 8-restart step.  The ~43 us this call costs over the device-resident step are the
  // copies' own latencies and the wake-up after the synchronise, not launch overhead.  Letting the kernels read theta from
  // and write the results to the pinned (device-mapped) host buffers directly, with no copy operations at all, was measured
  // too: 0.6307 vs 0.6326 ms at 8 restarts, 4.369 vs 4.379 ms at 64 -- within noise, and the generic path's sampler would
  // re-read theta over the bus once per MC sample, so the staged copies stay.)
  rc = sequence(st);
  if (rc != PRB_OK) return rc;
  if (synchronize) CUDA_TRY(cudaStreamSynchronize(st));
  return PRB_OK;
}

// ---- model lists ------------------------------------------------------------------------------------------------------
namespace {
struct MultiCtx {
  int nm;
  const Model* M[PRB_MAX_MODELS];
  Plan pl[PRB_MAX_MODELS];
  size_t base[PRB_MAX_MODELS];
  size_t total;
};

int multi_setup(prb_model* const* models, int n_models, int64_t n_points, int n_restarts, int d_theta, MultiCtx* c) {
  if (!models || n_models < 1 || n_models > PRB_MAX_MODELS) return PRB_ERR_INVALID;
  c->nm = n_models;
  size_t off = 0;
  for (int m = 0; m < n_models; ++m) {
    if (!models[m]) return PRB_ERR_INVALID;
    c->M[m] = reinterpret_cast<const Model*>(models[m]);
    if (c->M[m]->rff_m > 0) return PRB_ERR_INVALID;
    if (c->M[m]->d_k != c->M[0]->d_k || c->M[m]->n != c->M[0]->n) return PRB_ERR_INVALID;
    c->pl[m] = make_plan(*c->M[m], n_points, n_restarts > 0 ? n_restarts : 1, d_theta, 0);
    c->base[m] = off;
    off += (c->pl[m].total + 255) / 256 * 256;
  }
  c->total = off;
  return PRB_OK;
}

int fill_multi_args(const MultiCtx& c, const prb_multi_acq_desc* acq, void* ws, bool need_grad, MultiAcqArgs* ma) {
  if (!acq || (acq->kind != PRB_MACQ_CEI && acq->kind != PRB_MACQ_EHVI) || acq->n_models != c.nm) return PRB_ERR_INVALID;
  if (acq->kind == PRB_MACQ_CEI && (acq->objective_index < 0 || acq->objective_index >= c.nm)) return PRB_ERR_INVALID;
  if (acq->kind == PRB_MACQ_EHVI && (acq->n_cells < 0 || (acq->n_cells > 0 && acq->cell_bounds == nullptr) || c.nm < 2))
    return PRB_ERR_INVALID;
  memset(ma, 0, sizeof(MultiAcqArgs));
  ma->kind = acq->kind;
  ma->var_floor = acq->var_floor;
  ma->n_cells = acq->n_cells;
  ma->cell_bounds = acq->cell_bounds;
  ma->n_models = c.nm;
  ma->objective_index = acq->objective_index;
  ma->best_f = acq->best_f;
  ma->min_variance = acq->min_variance;
  ma->sigma_floor = acq->sigma_floor;
  for (int m = 0; m < c.nm; ++m) {
    char* w = static_cast<char*>(ws) + c.base[m];
    ma->has_lower[m] = acq->has_lower[m];
    ma->has_upper[m] = acq->has_upper[m];
    ma->lower[m] = acq->lower[m];
    ma->upper[m] = acq->upper[m];
    ma->mean_const[m] = c.M[m]->mean_const;
    ma->kss[m] = c.M[m]->kss;
    ma->norm_part[m] = at<double>(w, c.pl[m].off_norm);
    ma->mu_dot[m] = at<double>(w, c.pl[m].off_mu);
    ma->dmu[m] = need_grad ? at<double>(w, c.pl[m].off_dmu) : nullptr;
    ma->dvar[m] = need_grad ? at<double>(w, c.pl[m].off_dvar) : nullptr;
    ma->mtiles[m] = c.M[m]->Mp1 / GEMM_BM;
  }
  return PRB_OK;
}
}  // namespace

size_t prb_workspace_bytes_multi(prb_model* const* models, int32_t n_models, int64_t n_points, int32_t n_restarts,
                                 int32_t d_theta) {
  MultiCtx c;
  if (n_points < 0 || d_theta < 1 || multi_setup(models, n_models, n_points, n_restarts, d_theta, &c) != PRB_OK) return 0;
  return c.total;
}

int prb_mc_value_grad_multi(prb_model* const* models, int32_t n_models, const prb_space_desc* space,
                            const prb_multi_acq_desc* acq, const double* theta, int32_t b, int32_t mc,
                            const double* u_int, const double* u_cat, uint64_t seed, uint64_t offset, int32_t estimator,
                            double* ma_state, int32_t update_ma, const double* grad_out, double* out_value,
                            double* out_grad, double* out_acq_values, void* workspace, size_t workspace_bytes,
                            void* stream) {
  if (!theta || !out_value || b < 0 || mc < 1 || !workspace) return PRB_ERR_INVALID;
  if (grad_out && !out_grad) return PRB_ERR_INVALID;
  if (estimator != PRB_EST_REINFORCE && estimator != PRB_EST_REINFORCE_MA) return PRB_ERR_INVALID;
  if ((estimator == PRB_EST_REINFORCE_MA || update_ma) && !ma_state) return PRB_ERR_INVALID;
  DevSpace sp;
  int rc = make_space(space, &sp);
  if (rc != PRB_OK) return rc;
  if (b == 0) return PRB_OK;
  MultiCtx c;
  const int64_t B = int64_t(b) * mc;
  if ((rc = multi_setup(models, n_models, B, b, sp.d, &c)) != PRB_OK) return rc;
  if (sp.d_k != c.M[0]->d_k) return PRB_ERR_INVALID;
  if (workspace_bytes < c.total) return PRB_ERR_WORKSPACE;
  for (int m = 0; m < c.nm; ++m)
    if (B > c.pl[m].C) return PRB_ERR_UNSUPPORTED;
  const bool need_grad = grad_out != nullptr;
  const bool need_path = need_grad && sp.n_cont > 0;
  MultiAcqArgs ma;
  if ((rc = fill_multi_args(c, acq, workspace, need_path, &ma)) != PRB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* w0 = static_cast<char*>(workspace) + c.base[0];
  const Plan& p0 = c.pl[0];
  double* xk = at<double>(w0, p0.off_xk);
  uint8_t* z = at<uint8_t>(w0, p0.off_z);
  double* prob = at<double>(w0, p0.off_prob);
  double* a = out_acq_values ? out_acq_values : at<double>(w0, p0.off_a);
  const int ncols = int(B), ncols_pad = round_up(ncols, 128);
  CUDA_TRY(launch_pr_sample(sp, theta, b, mc, u_int, u_cat, seed, offset, xk, nullptr, need_grad ? z : nullptr,
                            need_grad ? prob : nullptr, nullptr, nullptr, st));
  GradGroup gq[PRB_MAX_MODELS];
  for (int m = 0; m < c.nm; ++m) {
    const Model& M = *c.M[m];
    char* w = static_cast<char*>(workspace) + c.base[m];
    const Plan& pl = c.pl[m];
    const bool gq_ok = need_path && build_grad_group(M.spec, cont_mask(sp), &gq[m]);
    if (need_path && !gq_ok) return PRB_ERR_UNSUPPORTED;
    CUDA_TRY(launch_kstar_panel(M, xk, ncols, ncols_pad, at<double>(w, pl.off_panelA), st, gq_ok ? &gq[m] : nullptr,
                                at<double>(w, pl.off_panelQ)));
    CUDA_TRY(launch_tma_lower(M, at<double>(w, pl.off_panelA), pl.C, ncols_pad,
                              need_path ? at<double>(w, pl.off_panelV) : nullptr, at<double>(w, pl.off_norm),
                              at<double>(w, pl.off_mu), nullptr, st));
  }
  CUDA_TRY(launch_multi_acq_epilogue(ma, ncols_pad, ncols, a, nullptr, nullptr, st));
  double* S0 = at<double>(w0, p0.off_S_part);
  if (need_path) {
    for (int m = 0; m < c.nm; ++m) {
      const Model& M = *c.M[m];
      char* w = static_cast<char*>(workspace) + c.base[m];
      const Plan& pl = c.pl[m];
      double* Sm = at<double>(w, pl.off_S_part);
      CUDA_TRY(launch_tma_upper_gradq(M, at<double>(w, pl.off_panelV), pl.C, ncols, ncols_pad, gq[m],
                                      at<double>(w, pl.off_panelQ), xk, ma.dmu[m], ma.dvar[m], Sm, st));
      if (m > 0) CUDA_TRY(launch_add_inplace(S0, Sm, size_t(M.Mp2 / GEMM_BM) * M.d_k * ncols_pad, st));
    }
  }
  DimMap dm;
  dm.n = need_path ? sp.n_cont : 0;
  for (int i = 0; i < sp.n_cont; ++i) dm.dims[i] = sp.cont_cols[i];
  const Model& M0 = *c.M[0];
  CUDA_TRY(launch_sep_step_finish(sp, dm, M0.d_k, 1, b, mc, a, z, prob, S0, M0.Mp2 / GEMM_BM, ncols_pad, estimator, ma_state,
                                  (update_ma && ma_state) ? 1 : 0, grad_out, out_value, out_grad, need_grad ? 1 : 0, nullptr,
                                  nullptr, nullptr, 0.0, nullptr, M0.done_ctr, st));
  return PRB_OK;
}

int prb_acq_points_multi(prb_model* const* models, int32_t n_models, const prb_multi_acq_desc* acq, const double* xk,
                         int64_t n_points, double* out_acq, double* out_mean, double* out_var, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (!xk || !out_acq || n_points < 0 || !workspace) return PRB_ERR_INVALID;
  if (n_points == 0) return PRB_OK;
  MultiCtx c;
  if (!models || n_models < 1 || !models[0]) return PRB_ERR_INVALID;
  int rc = multi_setup(models, n_models, n_points, 1, reinterpret_cast<const Model*>(models[0])->d_k, &c);
  if (rc != PRB_OK) return rc;
  if (workspace_bytes < c.total) return PRB_ERR_WORKSPACE;
  for (int m = 0; m < c.nm; ++m)
    if (n_points > c.pl[m].C) return PRB_ERR_UNSUPPORTED;
  MultiAcqArgs ma;
  if ((rc = fill_multi_args(c, acq, workspace, false, &ma)) != PRB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ncols = int(n_points), ncols_pad = round_up(ncols, 128);
  for (int m = 0; m < c.nm; ++m) {
    const Model& M = *c.M[m];
    char* w = static_cast<char*>(workspace) + c.base[m];
    const Plan& pl = c.pl[m];
    CUDA_TRY(launch_kstar_panel(M, xk, ncols, ncols_pad, at<double>(w, pl.off_panelA), st));
    CUDA_TRY(launch_tma_lower(M, at<double>(w, pl.off_panelA), pl.C, ncols_pad, nullptr, at<double>(w, pl.off_norm),
                              at<double>(w, pl.off_mu), nullptr, st));
  }
  CUDA_TRY(launch_multi_acq_epilogue(ma, ncols_pad, ncols, out_acq, out_mean, out_var, st));
  return PRB_OK;
}

int prb_mc_screen(prb_model* model, const prb_space_desc* space, const prb_acq_desc* acq, const double* theta, int32_t n_rows,
                  int32_t chunk_rows, int32_t mc, const double* u_int, const double* u_cat, uint64_t seed, uint64_t offset,
                  int32_t estimator, double* ma_state, double* out_value, void* workspace, size_t workspace_bytes,
                  int32_t chunk_cols, void* stream) {
  if (!model || !theta || !out_value || n_rows < 0 || chunk_rows < 1 || mc < 1 || !workspace) return PRB_ERR_INVALID;
  if (estimator != PRB_EST_REINFORCE && estimator != PRB_EST_REINFORCE_MA) return PRB_ERR_INVALID;
  const Model& m = *reinterpret_cast<Model*>(model);
  DevSpace sp;
  DevAcq aq;
  int rc = make_space(space, &sp);
  if (rc != PRB_OK) return rc;
  if ((rc = make_acq(acq, &aq)) != PRB_OK) return rc;
  if (sp.d_k != m.d_k) return PRB_ERR_INVALID;
  const int cmax = chunk_rows < n_rows ? chunk_rows : n_rows;
  if (cmax == 0) return PRB_OK;
  const Plan pl = make_plan(m, int64_t(cmax) * mc, cmax, sp.d, chunk_cols);
  if (workspace_bytes < pl.total) return PRB_ERR_WORKSPACE;
  for (int r0 = 0; r0 < n_rows; r0 += chunk_rows) {
    const int nr = n_rows - r0 < chunk_rows ? n_rows - r0 : chunk_rows;
    // every chunk is one forward call of the reference: value only, EMA state advanced once
    rc = mc_core(model, sp, aq, pl, workspace, theta + size_t(r0) * sp.d, nr, mc, u_int, u_cat, seed, offset, estimator,
                 ma_state, ma_state ? 1 : 0, nullptr, out_value + r0, nullptr, nullptr,
                 static_cast<cudaStream_t>(stream));
    if (rc != PRB_OK) return rc;
  }
  return PRB_OK;
}

int prb_analytic_value_grad(prb_model* model, const prb_space_desc* space, const prb_acq_desc* acq,
                            const double* theta, int32_t b, const int32_t* options, int32_t n_options,
                            const double* grad_out, double* out_value, double* out_grad, double* out_probs,
                            double* out_acq_values, void* workspace, size_t workspace_bytes, int32_t chunk_cols,
                            void* stream) {
  if (!model || !theta || !options || !out_value || b < 0 || n_options < 1 || !workspace) return PRB_ERR_INVALID;
  if (grad_out && !out_grad) return PRB_ERR_INVALID;
  const Model& m = *reinterpret_cast<Model*>(model);
  DevSpace sp;
  DevAcq aq;
  int rc = make_space(space, &sp);
  if (rc != PRB_OK) return rc;
  if ((rc = make_acq(acq, &aq)) != PRB_OK) return rc;
  if (sp.d_k != m.d_k) return PRB_ERR_INVALID;
  if (b == 0) return PRB_OK;
  const int64_t B = int64_t(b) * n_options;
  const Plan pl = make_plan(m, B, b, sp.d, chunk_cols);
  if (workspace_bytes < pl.total) return PRB_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* xk = at<double>(workspace, pl.off_xk);
  double* probs = out_probs ? out_probs : at<double>(workspace, pl.off_probs);
  double* a = out_acq_values ? out_acq_values : at<double>(workspace, pl.off_a);
  double* S_all = at<double>(workspace, pl.off_S_all);
  CUDA_TRY(launch_analytic_build(sp, theta, b, n_options, options, xk, probs, st));
  const bool need_path = grad_out != nullptr && sp.n_cont > 0;
  rc = run_engine(m, aq, pl, workspace, xk, B, need_path, cont_mask(sp), a, at<double>(workspace, pl.off_dmu),
                  at<double>(workspace, pl.off_dvar), S_all, nullptr, nullptr, st);
  if (rc != PRB_OK) return rc;
  CUDA_TRY(launch_analytic_reduce(sp, theta, b, n_options, options, a, probs, S_all, grad_out, out_value, out_grad, st));
  return PRB_OK;
}

int prb_analytic_adam(prb_model* model, const prb_space_desc* space, const prb_acq_desc* acq, double* theta, int32_t b,
                      const int32_t* options, int32_t n_options, const double* lower, const double* upper, double lr,
                      int32_t maxiter, int32_t step0, double* adam_state, double* out_value, void* workspace,
                      size_t workspace_bytes, int32_t chunk_cols, void* stream) {
  if (!model || !theta || !options || !lower || !upper || !adam_state || !out_value || b < 1 || n_options < 1 ||
      maxiter < 0 || !workspace)
    return PRB_ERR_INVALID;
  const Model& m = *reinterpret_cast<Model*>(model);
  DevSpace sp;
  DevAcq aq;
  int rc = make_space(space, &sp);
  if (rc != PRB_OK) return rc;
  if ((rc = make_acq(acq, &aq)) != PRB_OK) return rc;
  if (sp.d_k != m.d_k) return PRB_ERR_INVALID;
  const int64_t B = int64_t(b) * n_options;
  const Plan pl = make_plan(m, B, b, sp.d, chunk_cols);
  if (workspace_bytes < pl.total) return PRB_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* xk = at<double>(workspace, pl.off_xk);
  double* probs = at<double>(workspace, pl.off_probs);
  double* a = at<double>(workspace, pl.off_a);
  double* S_all = at<double>(workspace, pl.off_S_all);
  // One column panel and a GP model: a step is seven launches (six with a single row tile) -- clamp + enumeration, k* panel,
  // lower product [+ acquisition epilogue], upper product, gradient contraction, reduction + row-block sums + Adam.
  if (B <= pl.C && m.rff_m == 0 && !env_knobs().no_fused_step) {
    const int ncols = int(B), ncols_pad = round_up(int(B), 128), total = b * sp.d;
    double* panelA = at<double>(workspace, pl.off_panelA);
    double* panelV = at<double>(workspace, pl.off_panelV);
    double* panelQ = at<double>(workspace, pl.off_panelQ);
    double* norm = at<double>(workspace, pl.off_norm);
    double* mu = at<double>(workspace, pl.off_mu);
    double* S_part = at<double>(workspace, pl.off_S_part);
    double* dmu = at<double>(workspace, pl.off_dmu);
    double* dvar = at<double>(workspace, pl.off_dvar);
    double* Xc = at<double>(workspace, pl.off_Xc);
    double* grad = at<double>(workspace, pl.off_grad);
    double* ones = at<double>(workspace, pl.off_ones);
    double* val = at<double>(workspace, pl.off_val);
    int* step = at<int>(workspace, pl.off_step);
    unsigned int* done = m.done_ctr;
    CUDA_TRY(launch_fill(ones, 1.0, b, st));
    CUDA_TRY(cudaMemcpyAsync(step, &step0, sizeof(int), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(done, 0, sizeof(unsigned int), st));
    auto fused = [&](bool with_grad, double* value_out) -> int {
      const double* X = theta;
      if (with_grad) {  // X = clamp(theta) written to Xc by the enumeration kernel
        CUDA_TRY(launch_analytic_build(sp, theta, b, n_options, options, xk, probs, st, lower, upper, Xc));
        X = Xc;
      } else {  // final evaluation (gen.py:338-341): theta itself is clamped
        CUDA_TRY(launch_clamp(theta, lower, upper, b, sp.d, theta, st));
        CUDA_TRY(launch_analytic_build(sp, theta, b, n_options, options, xk, probs, st));
      }
      const bool need_path = with_grad && sp.n_cont > 0;
      GradGroup gq;
      const bool gq_ok = need_path && build_grad_group(m.spec, cont_mask(sp), &gq);
      CUDA_TRY(launch_kstar_panel(m, xk, ncols, ncols_pad, panelA, st, gq_ok ? &gq : nullptr, panelQ));
      FusedAcq fa;
      fa.aq = aq;
      fa.a = a;
      fa.dmu = dmu;
      fa.dvar = dvar;
      fa.ncols = ncols;
      const bool fuse = m.Mp1 == GEMM_BM;
      CUDA_TRY(launch_tma_lower(m, panelA, pl.C, ncols_pad, need_path ? panelV : nullptr, norm, mu, fuse ? &fa : nullptr,
                                st));
      if (!fuse)
        CUDA_TRY(launch_acq_epilogue(m, aq, norm, ncols_pad, mu, ncols, a, dmu, dvar, nullptr, nullptr, nullptr, 0, st));
      int jblocks = m.Mp2 / pl.grad_jrows;
      if (need_path && gq_ok) {
        CUDA_TRY(launch_tma_upper_gradq(m, panelV, pl.C, ncols, ncols_pad, gq, panelQ, xk, dmu, dvar, S_part, st));
        jblocks = m.Mp2 / GEMM_BM;
      } else if (need_path) {
        CUDA_TRY(launch_tma_upper(m, panelV, pl.C, ncols_pad, panelA, st));
        CUDA_TRY(launch_grad_contract(m, cont_mask(sp), xk, ncols, ncols_pad, panelA, dmu, dvar, S_part, pl.grad_jrows,
                                      st));
      }
      CUDA_TRY(launch_analytic_reduce(sp, X, b, n_options, options, a, probs, nullptr, with_grad ? ones : nullptr,
                                      value_out, with_grad ? grad : nullptr, st, need_path ? S_part : nullptr,
                                      jblocks, ncols_pad, with_grad ? theta : nullptr, adam_state,
                                      adam_state + total, lr, step, done));
      return PRB_OK;
    };
    rc = replay_steps(*reinterpret_cast<Model*>(model), 1, [&]() -> int { return fused(true, val); }, maxiter, st);
    if (rc != PRB_OK) return rc;
    return fused(false, out_value);
  }
  auto core = [&](const double* X, const double* go, double* value, double* g) -> int {
    CUDA_TRY(launch_analytic_build(sp, X, b, n_options, options, xk, probs, st));
    const bool need_path = go != nullptr && sp.n_cont > 0;
    int r = run_engine(m, aq, pl, workspace, xk, B, need_path, cont_mask(sp), a, at<double>(workspace, pl.off_dmu),
                       at<double>(workspace, pl.off_dvar), S_all, nullptr, nullptr, st);
    if (r != PRB_OK) return r;
    CUDA_TRY(launch_analytic_reduce(sp, X, b, n_options, options, a, probs, S_all, go, value, g, st));
    return PRB_OK;
  };
  return adam_graph_loop(*reinterpret_cast<Model*>(model), 1, core, theta, lower, upper, b, sp.d, lr, maxiter, step0, adam_state, out_value,
                         at<double>(workspace, pl.off_Xc), at<double>(workspace, pl.off_grad),
                         at<double>(workspace, pl.off_ones), at<double>(workspace, pl.off_val),
                         at<int>(workspace, pl.off_step), st);
}

int prb_mc_adam(prb_model* model, const prb_space_desc* space, const prb_acq_desc* acq, double* theta, int32_t b,
                int32_t mc, const double* u_int, const double* u_cat, uint64_t seed, uint64_t offset,
                int32_t estimator, double* ma_state, const double* lower, const double* upper, double lr,
                int32_t maxiter, int32_t step0, double* adam_state, double* out_value, void* workspace,
                size_t workspace_bytes, int32_t chunk_cols, void* stream) {
  if (!model || !theta || !lower || !upper || !adam_state || !out_value || b < 1 || mc < 1 || maxiter < 0 || !workspace)
    return PRB_ERR_INVALID;
  if (estimator != PRB_EST_REINFORCE && estimator != PRB_EST_REINFORCE_MA) return PRB_ERR_INVALID;
  if (estimator == PRB_EST_REINFORCE_MA && !ma_state) return PRB_ERR_INVALID;
  const Model& m = *reinterpret_cast<Model*>(model);
  DevSpace sp;
  DevAcq aq;
  int rc = make_space(space, &sp);
  if (rc != PRB_OK) return rc;
  if ((rc = make_acq(acq, &aq)) != PRB_OK) return rc;
  if (sp.d_k != m.d_k) return PRB_ERR_INVALID;
  const Plan pl = make_plan(m, int64_t(b) * mc, b, sp.d, chunk_cols);
  if (workspace_bytes < pl.total) return PRB_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* Xc = at<double>(workspace, pl.off_Xc);
  double* grad = at<double>(workspace, pl.off_grad);
  double* ones = at<double>(workspace, pl.off_ones);
  double* val = at<double>(workspace, pl.off_val);
  int* step = at<int>(workspace, pl.off_step);
  const int total = b * sp.d;
  CUDA_TRY(launch_fill(ones, 1.0, b, st));
  CUDA_TRY(cudaMemcpyAsync(step, &step0, sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(m.done_ctr, 0, sizeof(unsigned int), st));
  const int update_ma = ma_state ? 1 : 0;
  // separable structure on a single column panel: four fused launches per step instead of eleven
  BitMap bm;
  DimMap dm;
  const bool fused_step = pl.off_kc != 0 && sep_space_ok(m, sp, &bm, &dm) && int64_t(b) * mc <= pl.C &&
                          !(m.opt_dedup && m.sep_nbits <= DEDUP_MAX_BITS) && !env_knobs().no_fused_step;

  // experiment sizes (n + 1 <= 128, mc % 32 == 0): the whole step is ONE kernel (prb_small.cu)
  const bool sep_space = pl.off_kc != 0 && sep_space_ok(m, sp, &bm, &dm);
  bool small_gen = false;  // single-term kernels that do not separate use the generic variant of the same kernel
  bool small_step = fused_step && small_step_eligible(m, sp, dm, b, mc) &&
                    small_step_scratch_bytes(b, mc, sp.d) <= pl.small_bytes;
  if (!small_step && !sep_space && small_step_scratch_bytes(b, mc, sp.d) <= pl.small_bytes &&
      small_gen_eligible(m, sp, &dm, b, mc) && !env_knobs().no_fused_step)
    small_step = small_gen = true;
  if (small_step) CUDA_TRY(cudaMemsetAsync(at<char>(workspace, pl.off_small), 0, pl.small_bytes, st));
  if (small_step && small_adam_persistent_ok(b, mc)) {
    // the whole optimisation -- maxiter Adam steps and the final evaluation -- as one cooperative persistent kernel
    cudaError_t ce = launch_small_adam_persistent(m, small_gen, sp, aq, bm, dm, theta, lower, upper, b, mc, u_int, u_cat,
                                                  seed, offset, estimator, ma_state, update_ma, val, out_value, adam_state,
                                                  adam_state + total, lr, maxiter, step0, at<char>(workspace, pl.off_small),
                                                  st);
    if (ce == cudaSuccess) return PRB_OK;
    // fewer co-resident CTAs than the device attributes promise (MPS / SM partition): the replayed one-kernel step below
    if (ce != cudaErrorCooperativeLaunchTooLarge && ce != cudaErrorLaunchOutOfResources)
      return cuda_fail(ce, "launch_small_adam_persistent");
    (void)cudaGetLastError();
  }

  // kernel structures that do not separate: the six-launch fused step on a single column panel
  const bool generic_fused = !fused_step && !small_step && int64_t(b) * mc <= pl.C && !sep_space &&
                             !env_knobs().no_fused_step;

  auto one_iter = [&]() -> int {
    if (generic_fused)
      return generic_step(m, sp, aq, pl, workspace, theta, lower, upper, Xc, b, mc, u_int, u_cat, seed, offset, estimator,
                          ma_state, update_ma, true, nullptr, val, grad, theta, adam_state, adam_state + total, lr, step, st);
    if (small_step) {
      CUDA_TRY(launch_small_step(m, small_gen, sp, aq, bm, dm, theta, lower, upper, Xc, b, mc, u_int, u_cat, seed, offset, estimator,
                                 ma_state, update_ma, true, nullptr, val, grad, theta, adam_state, adam_state + total, lr,
                                 step, at<char>(workspace, pl.off_small), st));
      return PRB_OK;
    }
    if (fused_step)
      return sep_step(m, sp, aq, pl, workspace, bm, dm, theta, lower, upper, Xc, b, mc, u_int, u_cat, seed, offset,
                      estimator, ma_state, update_ma, true, nullptr, val, grad, theta, adam_state, adam_state + total,
                      lr, step, st);
    CUDA_TRY(launch_clamp(theta, lower, upper, b, sp.d, Xc, st));
    int r = mc_core(model, sp, aq, pl, workspace, Xc, b, mc, u_int, u_cat, seed, offset, estimator, ma_state,
                    update_ma, ones, val, grad, nullptr, st);
    if (r != PRB_OK) return r;
    CUDA_TRY(launch_adam_step(theta, grad, adam_state, adam_state + total, total, lr, step, st));
    return PRB_OK;
  };

  // ONE optimisation step is captured and replayed maxiter times (cached executable graph, see replay_steps)
  rc = replay_steps(*reinterpret_cast<Model*>(model), 0, one_iter, maxiter, st);
  if (rc != PRB_OK) return rc;
  // final clamp + value-only evaluation (gen.py:338-341); this call also advances the EMA state like any forward
  if (generic_fused)
    return generic_step(m, sp, aq, pl, workspace, theta, lower, upper, theta, b, mc, u_int, u_cat, seed, offset, estimator,
                        ma_state, update_ma, false, nullptr, out_value, nullptr, nullptr, nullptr, nullptr, lr, nullptr,
                        st);
  if (small_step) {
    CUDA_TRY(launch_small_step(m, small_gen, sp, aq, bm, dm, theta, lower, upper, theta, b, mc, u_int, u_cat, seed, offset, estimator,
                               ma_state, update_ma, false, nullptr, out_value, nullptr, nullptr, nullptr, nullptr, lr,
                               nullptr, at<char>(workspace, pl.off_small), st));
    return PRB_OK;
  }
  if (fused_step)
    return sep_step(m, sp, aq, pl, workspace, bm, dm, theta, lower, upper, theta, b, mc, u_int, u_cat, seed, offset,
                    estimator, ma_state, update_ma, false, nullptr, out_value, nullptr, nullptr, nullptr, nullptr, lr,
                    nullptr, st);
  CUDA_TRY(launch_clamp(theta, lower, upper, b, sp.d, theta, st));
  return mc_core(model, sp, aq, pl, workspace, theta, b, mc, u_int, u_cat, seed, offset, estimator, ma_state,
                 update_ma, nullptr, out_value, nullptr, nullptr, st);
}

}  // extern "C"
