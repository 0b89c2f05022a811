// b200sim.cu -- C ABI (include/b200sim.h) over the fused step kernel.
//
// Host responsibilities: validate the descriptor, derive the static tables the kernel
// wants (tree levels, children lists, per-link collidable-point lists, link inertia about
// the link origin), upload them once, pick the launch geometry, launch on the caller's
// stream.  No torch types cross this boundary.

#include "b200sim.h"
#include "b200sim_kernels.cuh"
#include "b200sim_rbda_kernels.cuh"
#include "b200sim_rigid_kernels.cuh"
#include "b200sim_step2.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <new>
#include <vector>

using namespace b200sim;

struct B200SimModel {
  int device = 0;
  int num_sms = 148;
  int max_smem_optin = 227 * 1024;
  int nL = 0, n = 0, nc = 0, depth = 0;
  int floating = 0, contact_model = 0, enable_friction = 1, flags = 0;
  double dt = 1e-3, g = -9.81, h_terrain = 0;
  double K = 1e6, D = 2e3, mu = 0.5, pexp = 0.5, qexp = 0.5;
  double tau_max = 3000, w_th = 30, w_max = 100;
  // host copies needed for update_link_params
  std::vector<double> cst_h;   // nL*CREC
  std::vector<double> csuc_h;  // nL*12
  std::vector<double> pt_h;    // nc*3
  std::vector<int> itab_h;
  int o_parent = 0, o_jtype = 0, o_lvl_start = 0, o_lvl_links = 0, o_child_start = 0, o_child_idx = 0,
      o_pt_start = 0, o_pt_idx = 0, o_pt_body = 0, o_pt_enabled = 0, o_anc = 0, o_ldepth = 0;
  int o_rows8 = 0, n_rows8 = 0, o_rows16 = 0, n_rows16 = 0;  // packed level-walk rows for G = 8 / 16 (0 rows: not available)
  int o_rows2_8 = 0, o_rows2_16 = 0;                          // the same rows in the format of step2_kernel
  // compact integer table of step2_kernel (only what it reads: a smaller block-static footprint buys environments)
  std::vector<int> itab2_h;
  int* itab2_d = nullptr;
  int t2_parent = 0, t2_jtype = 0, t2_pt_start = 0, t2_pt_idx = 0, t2_pt_body = 0, t2_pt_enabled = 0, t2_rows8 = 0, t2_rows16 = 0;
  double reg = 1e-6;
  double rx_tc = 0.02, rx_zeta = 1.0, rx_dmin = 0.9, rx_dmax = 0.95, rx_width = 1e-3, rx_mid = 0.5, rx_pow = 2.0;  // RelaxedRigid
  unsigned long long* dbg_d = nullptr;  // debug counters, allocated by b200sim_debug_counters
  // work lists of the rigid-contact cascade: one scratch buffer per (model, stream), so that steps of the same model on
  // different streams do not share counters
  struct RigidScratch {
    int* buf = nullptr; long long cap = 0;
    unsigned char* qp = nullptr; size_t qp_bytes = 0;  // contact-QP records of the split rigid level (one per work item)
    // split cascade: level 2 runs on a side stream next to the solve / resume launches of level 1
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  };
  std::unordered_map<void*, RigidScratch> rigid_scratch;
  std::mutex rigid_mutex;
  // stage buffers of b200sim_step_rk4, one per (model, stream) like the work lists above
  struct StageScratch { void* buf = nullptr; size_t bytes = 0; };
  std::unordered_map<void*, StageScratch> rk4_scratch;
  std::unordered_map<void*, StageScratch> vjp_scratch;  // dual inputs / outputs of b200sim_step_vjp
  // device blobs
  float *cst_f = nullptr, *csuc_f = nullptr, *pt_f = nullptr;
  double *cst_d = nullptr, *csuc_d = nullptr, *pt_d = nullptr;
  int* itab_d = nullptr;
  DualD *cst_dd = nullptr, *csuc_dd = nullptr, *pt_dd = nullptr;  // forward-mode AD blobs (value, tangent)
  bool cst_dd_has_tangent = false;                                // the resident cst_dd carries a mass direction
  // tuning
  int tune_G = 0, tune_epb = 0;
  int opt_flags = B200SIM_OPT_TMA_STORE;
};

namespace {

inline int max_threads_for(int G) { return G <= 8 ? B200SIM_KTHREADS : 512; }  // == LaunchBounds<G>::kThreads

#define CK(x)                         \
  do {                                \
    cudaError_t e_ = (x);             \
    if (e_ != cudaSuccess) return (int)e_; \
  } while (0)

template <typename T>
int upload(const std::vector<double>& h, T** d) {
  // padded to whole 16-byte chunks (+1 chunk) so the kernel can stage with 16-byte cp.async
  std::vector<T> tmp(((h.size() + 3) & ~size_t(3)) + 4, T(0));
  for (size_t i = 0; i < h.size(); ++i) tmp[i] = (T)h[i];
  if (!*d) CK(cudaMalloc((void**)d, tmp.size() * sizeof(T)));
  CK(cudaMemcpy(*d, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

void fill_link_consts(B200SimModel* m, const double* mass, const double* com, const double* inertia6) {
  for (int i = 0; i < m->nL; ++i) {
    double* c = m->cst_h.data() + (size_t)i * CREC;
    const double ms = mass[i];
    const double cx = com[3 * i], cy = com[3 * i + 1], cz = com[3 * i + 2];
    const double* I = inertia6 + 6 * i;  // xx xy xz yy yz zz at the CoM
    c[C_MASS] = ms;
    c[C_COM] = cx; c[C_COM + 1] = cy; c[C_COM + 2] = cz;
    // D = I_c + m S(c) S(c)^T = I_c + m (|c|^2 1 - c c^T)   (math/inertia.py:32-39)
    const double cc = cx * cx + cy * cy + cz * cz;
    c[C_DL + 0] = I[0] + ms * (cc - cx * cx);
    c[C_DL + 1] = I[1] - ms * cx * cy;
    c[C_DL + 2] = I[2] - ms * cx * cz;
    c[C_DL + 3] = I[3] + ms * (cc - cy * cy);
    c[C_DL + 4] = I[4] - ms * cy * cz;
    c[C_DL + 5] = I[5] + ms * (cc - cz * cz);
  }
}

bool is_identity4(const double* H, double tol = 0.0) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c)
      if (std::fabs(H[4 * r + c] - (r == c ? 1.0 : 0.0)) > tol) return false;
  return true;
}

size_t static_smem_bytes(const B200SimModel* m, size_t ts) {
  size_t w = (size_t)m->nL * CREC * ts;
  w += (((size_t)m->nc * 3 + 3) & ~size_t(3)) * ts;
  w += (((size_t)m->itab_h.size() + 3) & ~size_t(3)) * sizeof(int);
  return w;
}

size_t env_smem_bytes(const B200SimModel* m, size_t ts) {  // == env_ws_words<T>() * sizeof(T)
  size_t w = (size_t)m->nL * REC + (size_t)m->nc * PTREC;
  w = (w + 3) & ~size_t(3);
  return w * ts + 16;
}

struct Geometry {
  int G, epb, grid;
  size_t smem;
};

int pick_geometry(const B200SimModel* m, int dtype, long long B, Geometry* g, int force_G = 0) {
  const size_t ts = dtype == 2 ? 16 : (dtype == B200SIM_DTYPE_F64 ? 8 : 4);  // 2: Dual<double>
  const size_t st = static_smem_bytes(m, ts), pe = env_smem_bytes(m, ts);
  const size_t budget = (size_t)m->max_smem_optin - 1024;
  if (st + pe > budget) return B200SIM_E_TOO_LARGE;
  int G = force_G ? force_G : m->tune_G;
  if (G == 0) {
    // widest tree level / link count bound the useful lanes; 8 balances the sequential
    // level walks against the link-parallel phases for humanoid-size trees
    G = 8;
    // larger trees (ErgoCub-like, 50 links): 16 lanes halve the trips of the link-parallel phases and keep
    // the compact final phase available (nL <= 4 G + 1): measured 58.2 vs 68.1 us at batch 4096, 194 vs 232 us
    // at 16 384 (profiles/r01_lanes_ergocub.log)
    if (m->nL > 33) G = 16;
    while (G > 1 && G / 2 >= m->nL) G /= 2;
  }
  const int wg = 32 / G;  // groups per warp
  long long epb_smem = (long long)((budget - st) / pe);
  long long epb_thr = max_threads_for(G) / G;
  long long epb = std::min(epb_smem, epb_thr);
  if (m->tune_epb > 0) epb = std::min<long long>(epb, m->tune_epb);
  // spread the batch over all SMs in equally filled trips of the grid-stride loop (a last trip with a few
  // environments costs a full trip: 16 384 environments = 111 per SM run as 4 x 28, not 3 x 32 + 15)
  long long per_sm = (B + m->num_sms - 1) / m->num_sms;
  if (m->tune_epb == 0) {
    const long long trips = (per_sm + epb - 1) / std::max<long long>(epb, 1);
    epb = std::min(epb, std::max<long long>((per_sm + trips - 1) / std::max<long long>(trips, 1), 1));
  }
  // whole warps only
  epb = ((epb + wg - 1) / wg) * wg;
  while (epb > wg && (epb > epb_smem || epb > epb_thr)) epb -= wg;
  if (epb < 1 || epb > epb_smem) {
    // a single warp-group row does not fit next to the model: shrink to what fits
    epb = std::min<long long>(epb_smem, wg);
    if (epb < 1) return B200SIM_E_TOO_LARGE;
  }
  long long blocks = (B + epb - 1) / epb;
  const size_t smem = st + (size_t)epb * pe;
  long long resident = std::max<long long>(1, (long long)(budget + 1024) / (long long)(smem + 1024));
  long long cap = (long long)m->num_sms * resident;
  g->G = G;
  g->epb = (int)epb;
  g->grid = (int)std::max<long long>(1, std::min(blocks, cap));
  g->smem = smem;
  return 0;
}

template <typename T>
struct Blob;
template <>
struct Blob<float> {
  static const float* cst(const B200SimModel* m) { return m->cst_f; }
  static const float* csuc(const B200SimModel* m) { return m->csuc_f; }
  static const float* pt(const B200SimModel* m) { return m->pt_f; }
};
template <>
struct Blob<double> {
  static const double* cst(const B200SimModel* m) { return m->cst_d; }
  static const double* csuc(const B200SimModel* m) { return m->csuc_d; }
  static const double* pt(const B200SimModel* m) { return m->pt_d; }
};

template <>
struct Blob<DualD> {
  static const DualD* cst(const B200SimModel* m) { return m->cst_dd; }
  static const DualD* csuc(const B200SimModel* m) { return m->csuc_dd; }
  static const DualD* pt(const B200SimModel* m) { return m->pt_dd; }
};

template <typename T>
void fill_model_params(const B200SimModel* m, Params<T>& P) {
  P.cst = Blob<T>::cst(m);
  P.csuc = Blob<T>::csuc(m);
  P.pt_pos = Blob<T>::pt(m);
  P.itab = m->itab_d;
  P.itab_words = (int)m->itab_h.size();
  P.nL = m->nL; P.n = m->n; P.nc = m->nc; P.depth = m->depth;
  P.floating = m->floating; P.contact_model = m->contact_model; P.enable_friction = m->enable_friction;
  P.flags = m->flags | ((m->opt_flags & B200SIM_OPT_TMA_STORE) ? F_TMA_STORE : 0);
  P.o_parent = m->o_parent; P.o_jtype = m->o_jtype; P.o_lvl_start = m->o_lvl_start; P.o_lvl_links = m->o_lvl_links;
  P.o_child_start = m->o_child_start; P.o_child_idx = m->o_child_idx; P.o_pt_start = m->o_pt_start;
  P.o_pt_idx = m->o_pt_idx; P.o_pt_body = m->o_pt_body; P.o_pt_enabled = m->o_pt_enabled;
  P.o_anc = m->o_anc; P.o_ldepth = m->o_ldepth; P.reg = (T)m->reg;
  P.o_rows8 = m->o_rows8; P.n_rows8 = m->n_rows8; P.o_rows16 = m->o_rows16; P.n_rows16 = m->n_rows16;
  P.o_rows2_8 = m->o_rows2_8; P.o_rows2_16 = m->o_rows2_16;
  P.dbg = m->dbg_d;
  P.rx_tc = (T)m->rx_tc; P.rx_zeta = (T)m->rx_zeta; P.rx_dmin = (T)m->rx_dmin; P.rx_dmax = (T)m->rx_dmax;
  P.rx_width = (T)m->rx_width; P.rx_mid = (T)m->rx_mid; P.rx_pow = (T)m->rx_pow;
  P.dt = (T)m->dt; P.g = (T)m->g; P.h_terrain = (T)m->h_terrain;
  P.K = (T)m->K; P.D = (T)m->D; P.mu = (T)m->mu; P.pexp = (T)m->pexp; P.qexp = (T)m->qexp;
  P.tau_max = (T)m->tau_max; P.w_th = (T)m->w_th; P.w_max = (T)m->w_max;
}

template <typename T, int G, int SPEC = 0>
int launch_g(const Params<T>& P, const Geometry& g, cudaStream_t st, bool pdl = false) {
  auto kern = step_kernel<T, G, SPEC>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
  if (pdl) {
    // programmatic dependent launch: this launch may be set up while the previous kernel on the
    // stream still runs; the kernel orders its reads of the state with griddepcontrol.wait
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g.grid);
    cfg.blockDim = dim3(g.epb * G);
    cfg.dynamicSmemBytes = g.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kern, P);
  }
  kern<<<g.grid, g.epb * G, g.smem, st>>>(P);
  return (int)cudaGetLastError();
}

// The specialised instance (step_kernel<T, G, 1>) covers MODE_STEP of a floating-base model
// whose FK poses coincide with the ABA chain and whose successor transforms are identities,
// with soft contacts or no collidable points -- every BASELINE soft-contact configuration.
template <typename T>
bool specialised_step_applies(const B200SimModel* m, const Params<T>& P) {
  if (m->opt_flags & B200SIM_OPT_GENERIC_KERNEL) return false;
  if (P.mode != MODE_STEP || !P.floating) return false;
  if (P.flags & (F_SUC_NONID | F_GENERIC_FK)) return false;
  if (P.n > 0 && (m->n_rows8 == 0 || m->n_rows16 == 0)) return false;  // packed level-walk rows unavailable
  return P.nc == 0 || P.contact_model == 1;
}

// The second-generation kernel (b200sim_step2.cuh) serves what the specialised instance serves, for
// float / double, G = 8 / 16 lanes and trees with nL <= 4 G links.
size_t static_smem_bytes2(const B200SimModel* m, size_t ts) {
  size_t w = (size_t)m->nL * CREC * ts;
  w += (((size_t)m->nc * 3 + 3) & ~size_t(3)) * ts;
  w += (((size_t)m->itab2_h.size() + 3) & ~size_t(3)) * sizeof(int);
  return w;
}

int pick_geometry2(const B200SimModel* m, size_t ts, long long B, Geometry* g) {
  if (m->itab2_h.empty() || !m->itab2_d) return B200SIM_E_UNSUPPORTED;
  const size_t st = static_smem_bytes2(m, ts);
  // sharedMemPerBlockOptin bounds static + dynamic shared memory; 512 covers the kernel's static mbarrier array
  const size_t budget = (size_t)m->max_smem_optin - 512;
  int G = m->tune_G;
  if (G == 0) G = m->nL > 32 ? 16 : 8;
  if (G != 8 && G != 16) return B200SIM_E_UNSUPPORTED;
  if (m->nL > 4 * G) return B200SIM_E_UNSUPPORTED;
  if (((size_t)m->nL * 6 * ts) % 16 != 0) return B200SIM_E_UNSUPPORTED;  // (nL,6) rows of an environment leave / arrive as one bulk copy
  const size_t pe = step2_env_words(ts, m->nL, m->nc, G) * ts;
  if (st + pe > budget) return B200SIM_E_TOO_LARGE;
  const int wg = 32 / G;
  const long long epb_smem = (long long)((budget - st) / pe);
  const long long epb_thr = step2_max_threads(ts, G) / G;
  long long epb = std::min(epb_smem, epb_thr);
  if (m->tune_epb > 0) epb = std::min<long long>(epb, m->tune_epb);
  const long long per_sm = (B + m->num_sms - 1) / m->num_sms;
  if (m->tune_epb == 0) {  // equally filled trips of the grid-stride loop
    const long long trips = (per_sm + epb - 1) / std::max<long long>(epb, 1);
    epb = std::min(epb, std::max<long long>((per_sm + trips - 1) / std::max<long long>(trips, 1), 1));
  }
  epb = ((epb + wg - 1) / wg) * wg;  // whole warps
  while (epb > wg && (epb > epb_smem || epb > epb_thr)) epb -= wg;
  if (epb < 1 || epb > epb_smem || epb > epb_thr) return B200SIM_E_UNSUPPORTED;
  const long long blocks = (B + epb - 1) / epb;
  g->G = G;
  g->epb = (int)epb;
  g->grid = (int)std::max<long long>(1, std::min<long long>(blocks, m->num_sms));
  g->smem = st + (size_t)epb * pe;
  return 0;
}

template <typename T>
int try_launch_step2(const B200SimModel*, Params<T>&, void*, bool* done) { *done = false; return 0; }

template <typename T>
int launch_step2_impl(const B200SimModel* m, Params<T>& P, void* stream, bool* done) {
  *done = false;
  if (m->opt_flags & B200SIM_OPT_STEP_V1) return 0;
  Geometry g;
  if (pick_geometry2(m, sizeof(T), P.B, &g) != 0) return 0;
  // the cache outputs leave with cp.async.bulk: 16-byte aligned leaves
  if (((uintptr_t)P.iXl | (uintptr_t)P.W_H_L | (uintptr_t)P.W_v) % 16 != 0) return 0;
  *done = true;
  P.envs_per_block = g.epb;
  P.ws2_words = (int)step2_env_words(sizeof(T), m->nL, m->nc, g.G);
  // the kernel stages the compact integer table
  P.itab = m->itab2_d;
  P.itab_words = (int)m->itab2_h.size();
  P.o_parent = m->t2_parent; P.o_jtype = m->t2_jtype; P.o_pt_start = m->t2_pt_start; P.o_pt_idx = m->t2_pt_idx;
  P.o_pt_body = m->t2_pt_body; P.o_pt_enabled = m->t2_pt_enabled; P.o_rows2_8 = m->t2_rows8; P.o_rows2_16 = m->t2_rows16;
  // cp.async.bulk needs 16-byte aligned, 16-byte granular rows per environment
  P.flags &= ~F_BULK_IN;
  if (!(m->opt_flags & B200SIM_OPT_NO_BULK_IN) && P.Hin && P.Vin && ((uintptr_t)P.Hin % 16 == 0) && ((uintptr_t)P.Vin % 16 == 0) &&
      (((size_t)m->nL * 6 * sizeof(T)) % 16 == 0))
    P.flags |= F_BULK_IN;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev != m->device) CK(cudaSetDevice(m->device));
  const int rc = launch_step2<T>(P, g.G, g.grid, g.epb * g.G, g.smem, (cudaStream_t)stream, !(m->opt_flags & B200SIM_OPT_NO_PDL));
  if (dev != m->device) cudaSetDevice(dev);
  return rc;
}
template <>
int try_launch_step2<float>(const B200SimModel* m, Params<float>& P, void* stream, bool* done) { return launch_step2_impl(m, P, stream, done); }
template <>
int try_launch_step2<double>(const B200SimModel* m, Params<double>& P, void* stream, bool* done) { return launch_step2_impl(m, P, stream, done); }

template <typename T>
int launch(const B200SimModel* m, Params<T>& P, int dtype, void* stream) {
  if (specialised_step_applies(m, P) && !P.over_count && !P.work_count) {
    bool done = false;
    const int rc2 = try_launch_step2<T>(m, P, stream, &done);
    if (done) return rc2;
  }
  Geometry g;
  int rc = pick_geometry(m, dtype, P.B, &g);
  if (rc) return rc;
  P.envs_per_block = g.epb;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev != m->device) CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  const bool spec = specialised_step_applies(m, P);
  const bool pdl = spec && !(m->opt_flags & B200SIM_OPT_NO_PDL);
  switch (g.G) {
    case 1: rc = launch_g<T, 1>(P, g, st); break;
    case 2: rc = launch_g<T, 2>(P, g, st); break;
    case 4: rc = launch_g<T, 4>(P, g, st); break;
    case 8: rc = spec ? launch_g<T, 8, 1>(P, g, st, pdl) : launch_g<T, 8>(P, g, st); break;
    case 16: rc = spec ? launch_g<T, 16, 1>(P, g, st, pdl) : launch_g<T, 16>(P, g, st); break;
    case 32: rc = launch_g<T, 32>(P, g, st); break;
    default: rc = B200SIM_E_INVALID;
  }
  if (dev != m->device) cudaSetDevice(dev);
  return rc;
}

// ---- rigid contacts: a cascade of three launches on the caller's stream
//   level 0: the fused step kernel (G lanes per environment) finishes every environment whose
//            collidable points stay above the ground; the others land on work list 1;
//   level 1: the rigid kernel, one warp per listed environment, shared-memory workspace sized
//            for RIGID_CAP1 simultaneously active points; environments with more -> list 2;
//   level 2: the rigid kernel with a full-size workspace (only if nc > RIGID_CAP1).
constexpr int RIGID_CAP1 = 12;  // 12 active points: 7 resident warps per SM instead of 5 at 16 (the rest overflows to the full-size level)

template <typename T, typename S>
int rigid_geometry(const B200SimModel* m, long long B, int cap, int qp_mode, int* warps, int* grid, size_t* smem) {
  const size_t st = static_smem_bytes(m, sizeof(T));
  const RigidLayout L = rigid_layout<T, S>(m->nL, m->nc, m->depth, cap, qp_mode);
  const size_t budget = (size_t)m->max_smem_optin - 1024;
  if (st + L.total > budget) return B200SIM_E_TOO_LARGE;
  long long w = std::min<long long>(RIGID_MAX_WARPS, (long long)((budget - st) / L.total));
  const long long per_sm = (B + m->num_sms - 1) / m->num_sms;
  w = std::min(w, std::max<long long>(per_sm, 1));
  *warps = (int)w;
  *smem = st + (size_t)w * L.total;
  const long long blocks_per_sm = std::max<long long>(1, std::min<long long>(4, (long long)((size_t)228 * 1024 / (*smem + 1024))));
  const long long want = (B + w - 1) / w;
  *grid = (int)std::min<long long>(want, (long long)m->num_sms * blocks_per_sm);
  return 0;
}

template <typename T, typename S>
int launch_rigid_level(const B200SimModel* m, Params<T>& P, int cap, cudaStream_t st) {
  int warps = 0, grid = 0;
  size_t smem = 0;
  int rc = rigid_geometry<T, S>(m, P.B, cap, m->contact_model == B200SIM_CONTACT_RELAXED_RIGID ? 2 : P.qp_mode, &warps, &grid, &smem);
  if (rc) return rc;
  P.envs_per_block = warps;
  P.na_cap = cap;
  auto kern = rigid_step_kernel<T, S>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, warps * 32, smem, st>>>(P);
  return (int)cudaGetLastError();
}

// work lists: [0..7] counters, then three lists of `cap` ints.  Grown on demand (not during a
// stream capture: run one eager step of the largest batch first).
//   counters: [0] list 1 (level 0 -> level 1), [1] list 2 (level 1 -> level 2), [2] next item of the solve launch,
//             [3] list 3 (split cascade: impact overflows of the resuming launch -> second level-2 launch)
constexpr int RIGID_COUNTERS = 8;
int ensure_rigid_scratch(B200SimModel* m, long long B, cudaStream_t st, B200SimModel::RigidScratch** out) {
  std::lock_guard<std::mutex> lock(m->rigid_mutex);
  B200SimModel::RigidScratch& sc = m->rigid_scratch[(void*)st];
  if (!sc.buf || sc.cap < B || !sc.aux) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone)
      return B200SIM_E_UNSUPPORTED;  // cannot allocate inside a capture: run one eager step of the largest batch first
    if (!sc.aux) {  // all three or none: a failure must not leave a stream without its events behind
      cudaStream_t aux = nullptr;
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      cudaError_t err = cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking);
      if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e0, cudaEventDisableTiming);
      if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e1, cudaEventDisableTiming);
      if (err != cudaSuccess) {
        if (e1) cudaEventDestroy(e1);
        if (e0) cudaEventDestroy(e0);
        if (aux) cudaStreamDestroy(aux);
        return (int)err;
      }
      sc.aux = aux; sc.ev_fork = e0; sc.ev_join = e1;
    }
    if (!sc.buf || sc.cap < B) {
      if (sc.buf) {
        CK(cudaStreamSynchronize(st));
        CK(cudaFree(sc.buf));
      }
      sc.buf = nullptr;
      sc.cap = 0;
      CK(cudaMalloc((void**)&sc.buf, sizeof(int) * (RIGID_COUNTERS + 3 * (size_t)B)));
      sc.cap = B;
    }
  }
  *out = &sc;
  return 0;
}

// records of the split rigid level, grown like the work lists (same rule under stream capture)
int ensure_qp_scratch(B200SimModel* m, size_t bytes, cudaStream_t st, unsigned char** buf) {
  std::lock_guard<std::mutex> lock(m->rigid_mutex);
  B200SimModel::RigidScratch& sc = m->rigid_scratch[(void*)st];
  if (!sc.qp || sc.qp_bytes < bytes) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) return B200SIM_E_UNSUPPORTED;
    if (sc.qp) {
      CK(cudaStreamSynchronize(st));
      CK(cudaFree(sc.qp));
    }
    sc.qp = nullptr;
    sc.qp_bytes = 0;
    CK(cudaMalloc((void**)&sc.qp, bytes));
    sc.qp_bytes = bytes;
  }
  *buf = sc.qp;
  return 0;
}

// the contact QPs of a split level's work items (rigid_qp_kernel)
template <typename S>
int launch_rigid_qp(const B200SimModel* m, long long B, const int* work_count, int* next_item, unsigned char* qp_buf, long long qp_stride,
                    int cap, double mu, double tol, int* status, unsigned long long* dbg, cudaStream_t st) {
  const QpLayout L = qp_layout<S>(cap);
  const size_t smem = (size_t)QP_WARPS * L.total;
  static const int minb = [] { const char* e = std::getenv("B200SIM_QP_MINB"); return e ? std::atoi(e) : 8; }();  // diagnostic A/B
  auto kern = minb == 10 ? rigid_qp_kernel<S, 10> : rigid_qp_kernel<S, 8>;  // 128 registers without spills beat 96 with (4.64 / 4.82 ms)
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;  // exactly one resident wave: the blocks draw their items from `next_item`
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * QP_WARPS, smem));
  const long long want = (B + QP_WARPS - 1) / QP_WARPS;
  const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)m->num_sms * std::max(per_sm, 1)));
  kern<<<grid, 32 * QP_WARPS, smem, st>>>(work_count, next_item, qp_buf, qp_stride, cap, (S)mu, (S)tol, status, dbg);
  return (int)cudaGetLastError();
}

// RigidContacts, split cascade.  Level 1 is three launches -- assemble (rigid kernel, qp_mode 1), solve
// (rigid_qp_kernel: 3-4x the resident warps of the monolithic kernel, where the interior-point iteration is 88 % of the
// time), resume (qp_mode 2).  Level 2 (environments with more active points than level 1 holds, monolithic full-size
// kernel, two warps per SM) starts on a side stream as soon as the assembling launch has listed them and runs NEXT TO
// the solve / resume launches; the few environments whose IMPACT overflows level 1 go to a second list and a second,
// usually empty, level-2 launch after the join.
template <typename T, typename S>
int launch_rigid_split(B200SimModel* m, const Params<T>& P, B200SimModel::RigidScratch* sc, int cap1, cudaStream_t st) {
  int* cnt = sc->buf;
  int* list1 = cnt + RIGID_COUNTERS;
  int* list2 = list1 + sc->cap;
  int* list3 = list2 + sc->cap;
  const size_t stride = qp_record_bytes<T, S>(m->nL, m->nc, cap1);
  unsigned char* qp = nullptr;
  int rc = ensure_qp_scratch(m, stride * (size_t)P.B, st, &qp);
  if (rc) return rc;
  const bool level2 = cap1 < m->nc;
  Params<T> P1 = P;
  P1.work_count = cnt; P1.work_list = list1;
  P1.over_count = cnt + 1; P1.over_list = list2;
  P1.qp_buf = qp; P1.qp_stride = (long long)stride;
  P1.qp_mode = 1;
  rc = launch_rigid_level<T, S>(m, P1, cap1, st);
  if (rc) return rc;
  if (level2) {
    CK(cudaEventRecord(sc->ev_fork, st));
    CK(cudaStreamWaitEvent(sc->aux, sc->ev_fork, 0));
    Params<T> P2 = P;
    P2.work_count = cnt + 1; P2.work_list = list2;
    rc = launch_rigid_level<T, S>(m, P2, m->nc, sc->aux);
    // the join is recorded whatever happened, so that a capture of `st` never ends with the side stream forked
    CK(cudaEventRecord(sc->ev_join, sc->aux));
  }
  // same stopping rule as the monolithic kernel: a float64 solve of float32 data stops at the resolution of the data
  double tol = (sizeof(S) == 8) ? (sizeof(T) == 4 ? 1e-9 : 1e-12) : 1e-5;
  static const double tol_f32 = [] { const char* e = std::getenv("B200SIM_QP_TOL_F32"); return e ? std::atof(e) : 0.0; }();  // diagnostic A/B
  if (tol_f32 > 0 && sizeof(S) == 8 && sizeof(T) == 4) tol = tol_f32;
  if (!rc) rc = launch_rigid_qp<S>(m, P.B, cnt, cnt + 2, qp, (long long)stride, cap1, (double)P.mu, tol, P.status, P.dbg, st);
  P1.qp_mode = 2;
  P1.over_count = cnt + 3; P1.over_list = list3;
  if (!rc) rc = launch_rigid_level<T, S>(m, P1, cap1, st);
  if (level2) {
    CK(cudaStreamWaitEvent(st, sc->ev_join, 0));
    if (!rc) {
      Params<T> P3 = P;
      P3.work_count = cnt + 3; P3.work_list = list3;
      rc = launch_rigid_level<T, S>(m, P3, m->nc, st);
    }
  }
  return rc;
}

template <typename T>
int launch_rigid(const B200SimModel* cm, Params<T>& P, int dtype, void* stream) {
  B200SimModel* m = const_cast<B200SimModel*>(cm);
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev != m->device) CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  B200SimModel::RigidScratch* sc = nullptr;
  int rc = ensure_rigid_scratch(m, P.B, st, &sc);
  int* cnt = sc ? sc->buf : nullptr;
  int* list1 = cnt ? cnt + RIGID_COUNTERS : nullptr;
  int* list2 = cnt ? list1 + sc->cap : nullptr;
  if (!rc) rc = (int)cudaMemsetAsync(cnt, 0, RIGID_COUNTERS * sizeof(int), st);
  if (!rc) {
    // level 0: never reads the cached link velocities (they may be the pre-impact ones)
    Params<T> P0 = P;
    P0.Hin = nullptr; P0.Vin = nullptr;
    P0.over_count = cnt; P0.over_list = list1;
    rc = launch(m, P0, dtype, stream);
  }
  const bool qp32 = sizeof(T) == 4 && (m->opt_flags & B200SIM_OPT_RIGID_QP_F32);
  const int cap1 = std::min(m->nc, RIGID_CAP1);
  // RigidContacts: split cascade (the records of a very large batch would not be worth their memory: monolithic then)
  const bool split = m->contact_model == B200SIM_CONTACT_RIGID && !(m->opt_flags & B200SIM_OPT_RIGID_MONO) &&
                     (size_t)P.B * qp_record_bytes<double, double>(m->nL, m->nc, cap1) <= ((size_t)2 << 30);
  if (!rc && split) {
    rc = qp32 ? launch_rigid_split<T, T>(m, P, sc, cap1, st) : launch_rigid_split<T, double>(m, P, sc, cap1, st);
  } else {
    if (!rc) {
      Params<T> P1 = P;
      P1.work_count = cnt; P1.work_list = list1;
      P1.over_count = cnt + 1; P1.over_list = list2;
      rc = qp32 ? launch_rigid_level<T, T>(m, P1, cap1, st) : launch_rigid_level<T, double>(m, P1, cap1, st);
    }
    if (!rc && cap1 < m->nc) {
      Params<T> P2 = P;
      P2.work_count = cnt + 1; P2.work_list = list2;
      rc = qp32 ? launch_rigid_level<T, T>(m, P2, m->nc, st) : launch_rigid_level<T, double>(m, P2, m->nc, st);
    }
  }
  if (dev != m->device) cudaSetDevice(dev);
  return rc;
}

enum RbdaKind { RBDA_RNEA = 0, RBDA_CRBA = 1 };

template <typename T, int G>
int launch_rbda_g(int kind, const Params<T>& P, const RbdaArgs<T>& A, const Geometry& g, cudaStream_t st) {
  if (kind == RBDA_RNEA) {
    auto kern = rnea_kernel<T, G>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    kern<<<g.grid, g.epb * G, g.smem, st>>>(P, A);
  } else {
    auto kern = crba_kernel<T, G>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    kern<<<g.grid, g.epb * G, g.smem, st>>>(P, A);
  }
  return (int)cudaGetLastError();
}

template <typename T>
int launch_rbda(const B200SimModel* m, int kind, Params<T>& P, const RbdaArgs<T>& A, int dtype, void* stream) {
  Geometry g;
  int rc = pick_geometry(m, dtype, P.B, &g);
  if (rc) return rc;
  P.envs_per_block = g.epb;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev != m->device) CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (kind == RBDA_CRBA) {
    const size_t N = 6 + (size_t)m->n;
    rc = (int)cudaMemsetAsync(A.M, 0, (size_t)P.B * N * N * sizeof(T), st);
  }
  if (!rc) {
    switch (g.G) {
      case 1: rc = launch_rbda_g<T, 1>(kind, P, A, g, st); break;
      case 2: rc = launch_rbda_g<T, 2>(kind, P, A, g, st); break;
      case 4: rc = launch_rbda_g<T, 4>(kind, P, A, g, st); break;
      case 8: rc = launch_rbda_g<T, 8>(kind, P, A, g, st); break;
      case 16: rc = launch_rbda_g<T, 16>(kind, P, A, g, st); break;
      case 32: rc = launch_rbda_g<T, 32>(kind, P, A, g, st); break;
      default: rc = B200SIM_E_INVALID;
    }
  }
  if (dev != m->device) cudaSetDevice(dev);
  return rc;
}

template <typename T>
int rnea_t(const B200SimModel* m, int dtype, int64_t B, const void* s, const void* sd, const void* q, const void* vlin,
           const void* omega, const void* p, const void* avd, const void* sdd, const void* fext, void* W_f0, void* tau,
           void* stream) {
  Params<T> P;
  std::memset(&P, 0, sizeof(P));
  fill_model_params(m, P);
  P.B = B;
  P.s = (const T*)s; P.sd = (const T*)sd; P.q = (const T*)q; P.vlin = (const T*)vlin; P.omega = (const T*)omega;
  P.p = (const T*)p; P.fext = (const T*)fext;
  P.nsteps = 1;
  RbdaArgs<T> A;
  std::memset(&A, 0, sizeof(A));
  A.avd = (const T*)avd; A.sdd = (const T*)sdd; A.W_f0 = (T*)W_f0; A.tau_o = (T*)tau;
  return launch_rbda(m, RBDA_RNEA, P, A, dtype, stream);
}

template <typename T>
int crba_t(const B200SimModel* m, int dtype, int64_t B, const void* s, void* M, void* stream) {
  Params<T> P;
  std::memset(&P, 0, sizeof(P));
  fill_model_params(m, P);
  P.B = B;
  P.s = (const T*)s;
  P.nsteps = 1;
  RbdaArgs<T> A;
  std::memset(&A, 0, sizeof(A));
  A.M = (T*)M;
  return launch_rbda(m, RBDA_CRBA, P, A, dtype, stream);
}

template <typename T>
int step_t(const B200SimModel* m, int dtype, int64_t B, const void* s, const void* sd, const void* q,
           const void* vlin, const void* omega, const void* p, const void* mt, const void* tau, const void* fext,
           void* s_o, void* sd_o, void* q_o, void* vlin_o, void* omega_o, void* p_o, void* m_o, void* W_H_B,
           void* iXl, void* W_H_L, void* W_v, int nsteps, long long tau_stride, long long fext_stride, const void* Hin,
           const void* Vin, void* stream, int* status = nullptr, int fext_repr = 0) {
  Params<T> P;
  std::memset(&P, 0, sizeof(P));
  fill_model_params(m, P);
  P.status = status;
  P.fext_repr = fext_repr;
  P.B = B;
  P.s = (const T*)s; P.sd = (const T*)sd; P.q = (const T*)q; P.vlin = (const T*)vlin; P.omega = (const T*)omega;
  P.p = (const T*)p; P.m = (const T*)mt; P.tau = (const T*)tau; P.fext = (const T*)fext;
  P.s_o = (T*)s_o; P.sd_o = (T*)sd_o; P.q_o = (T*)q_o; P.vlin_o = (T*)vlin_o; P.omega_o = (T*)omega_o;
  P.p_o = (T*)p_o; P.m_o = (T*)m_o;
  P.W_H_B = (T*)W_H_B; P.iXl = (T*)iXl; P.W_H_L = (T*)W_H_L; P.W_v = (T*)W_v;
  P.nsteps = nsteps; P.tau_step_stride = tau_stride; P.fext_step_stride = fext_stride;
  P.Hin = (const T*)Hin; P.Vin = (const T*)Vin;
  // cp.async.bulk needs 16-byte aligned, 16-byte granular blocks per environment
  if ((m->opt_flags & B200SIM_OPT_BULK_IN) && Hin && Vin && ((uintptr_t)Hin % 16 == 0) && ((uintptr_t)Vin % 16 == 0) && (((size_t)m->nL * 6 * sizeof(T)) % 16 == 0))
    P.flags |= F_BULK_IN;
  P.mode = MODE_STEP;
  if (m->contact_model >= B200SIM_CONTACT_RIGID && m->nc > 0) {
    if (nsteps == 1) return launch_rigid(m, P, dtype, stream);
    // n consecutive rigid steps = n cascades on the stream.  From the second step on the state is stepped IN PLACE in
    // the output buffers and the cached link transforms / velocities written by the previous step are the cached inputs
    // -- exactly what repeated `step` calls hand over, including the pre-impact link velocities the reference leaves in
    // its cache (rbda/contacts/rigid.py:429-434).
    if (!W_H_L || !W_v) return B200SIM_E_UNSUPPORTED;
    Params<T> Pk = P;
    Pk.nsteps = 1;
    Pk.tau_step_stride = 0;
    Pk.fext_step_stride = 0;
    for (int k = 0; k < nsteps; ++k) {
      Pk.tau = P.tau ? P.tau + (long long)k * tau_stride : nullptr;
      Pk.fext = P.fext ? P.fext + (long long)k * fext_stride : nullptr;
      if (k == 1) {
        Pk.s = P.s_o; Pk.sd = P.sd_o; Pk.q = P.q_o; Pk.vlin = P.vlin_o; Pk.omega = P.omega_o; Pk.p = P.p_o;
        Pk.Hin = P.W_H_L; Pk.Vin = P.W_v;
      }
      const int rc = launch_rigid(m, Pk, dtype, stream);
      if (rc) return rc;
    }
    return 0;
  }
  return launch(m, P, dtype, stream);
}

template <typename T>
int fk_t(const B200SimModel* m, int dtype, int64_t B, const void* s, const void* sd, const void* q, const void* vlin,
         const void* omega, const void* p, void* q_o, void* W_H_B, void* iXl, void* W_H_L, void* W_v, void* stream) {
  Params<T> P;
  std::memset(&P, 0, sizeof(P));
  fill_model_params(m, P);
  P.B = B;
  P.s = (const T*)s; P.sd = (const T*)sd; P.q = (const T*)q; P.vlin = (const T*)vlin; P.omega = (const T*)omega;
  P.p = (const T*)p;
  P.q_o = (T*)q_o;
  P.W_H_B = (T*)W_H_B; P.iXl = (T*)iXl; P.W_H_L = (T*)W_H_L; P.W_v = (T*)W_v;
  P.nsteps = 1;
  P.mode = MODE_FK;
  return launch(m, P, dtype, stream);
}

template <typename T>
int aba_t(const B200SimModel* m, int dtype, int64_t B, const void* s, const void* sd, const void* q, const void* vlin,
          const void* omega, const void* p, const void* tau, const void* fext, void* avd, void* sdd, void* stream) {
  Params<T> P;
  std::memset(&P, 0, sizeof(P));
  fill_model_params(m, P);
  P.B = B;
  P.s = (const T*)s; P.sd = (const T*)sd; P.q = (const T*)q; P.vlin = (const T*)vlin; P.omega = (const T*)omega;
  P.p = (const T*)p; P.tau = (const T*)tau; P.fext = (const T*)fext;
  P.avd = (T*)avd; P.sdd_o = (T*)sdd;
  P.nsteps = 1;
  P.mode = MODE_ABA;
  return launch(m, P, dtype, stream);
}

// ---- RungeKutta4 (api/integrators.py:91-156) --------------------------------------------------------------------
// Four evaluations of the system dynamics (the MODE_DYN launch of the step kernel: contacts -> ABA -> position dynamics)
// joined by one small elementwise launch each, all on the caller's stream with no host round trip; the stage state, the
// running sum of the slopes and the resultant torques live in a per-(model, stream) scratch block.
template <typename T>
struct Rk4Bufs {
  // stage state (what the dynamics launch reads)
  T *s, *sd, *q, *vl, *om, *p, *m;
  // slopes written by the dynamics launch (the slope of s is the stage's sd)
  T *pd, *qd, *vd, *sdd, *md;
  // running sums of the slopes, the normalised input quaternion and the resultant torques
  T *ks, *ksd, *kq, *kvl, *kom, *kp, *km, *q0, *tau;
};

template <typename T>
struct Rk4Io {
  const T *s, *sd, *q, *vl, *om, *p, *m, *tau_ref;
  T *s_o, *sd_o, *q_o, *vl_o, *om_o, *p_o, *m_o;
};

// stage -1: x_stage = x0 (quaternion normalised, integrators.py:105-110), sums = 0, tau = actuation model
// (api/actuation_model.py:7-126, evaluated ONCE on the input state: api/model.py:2658).
// stage 0..2: sums += w k; x_stage = x0 + c k.   stage 3: out = x0 + dt/6 (sums + k).
template <typename T>
__global__ void rk4_stage_kernel(long long B, int n, int nc, int stage, T dt, Rk4Io<T> io, Rk4Bufs<T> b, const T* __restrict__ cst,
                                 int enable_friction, T tau_max, T w_th, T w_max) {
  const int E = n + 13 + 3 * nc;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= B * E) return;
  const long long e = tid / E;
  const int k = (int)(tid - e * E);
  const T w = (stage == 0) ? T(1) : T(2);
  const T c = (stage == 2) ? dt : T(0.5) * dt;
  // one (x0, slope, sum, stage, out) tuple per element; joints carry two (s and sd) so that the slope of s -- the
  // stage's own sd -- is read before it is overwritten
  auto advance = [&](T x0, T slope, T* sum, T* xs, T* out) {
    if (stage < 0) { *sum = T(0); *xs = x0; return; }
    if (stage < 3) { *sum += w * slope; *xs = x0 + c * slope; return; }
    *out = x0 + dt * ((*sum + slope) / T(6));
  };
  if (k < n) {
    const long long j = e * n + k;
    const T s0 = io.s[j], sd0 = io.sd[j];
    if (stage < 0) {
      const T* cl = cst + (size_t)(k + 1) * CREC;
      const T lower = min_t(s0 - cl[C_SMIN], T(0));
      const T upper = max_t(s0 - cl[C_SMAX], T(0));
      T tlim = -cl[C_KS] * (lower + upper);
      tlim = tlim - tlim * cl[C_KD] * sd0;
      T tfr = T(0);
      if (enable_friction) {
        const T sg = (sd0 > T(0)) ? T(1) : ((sd0 < T(0)) ? T(-1) : T(0));
        tfr = -(cl[C_KC] * sg + cl[C_KV] * sd0);
      }
      const T tt = (io.tau_ref ? io.tau_ref[j] : T(0)) + tfr + tlim;
      const T av = abs_t(sd0);
      T lim;
      if (av <= w_th) lim = tau_max;
      else if (av <= w_max) lim = tau_max * (T(1) - (av - w_th) / (w_max - w_th));
      else lim = T(0);
      b.tau[j] = min_t(max_t(tt, -lim), lim);
    }
    const T sd_stage = (stage < 0) ? T(0) : b.sd[j];
    const T sdd = (stage < 0) ? T(0) : b.sdd[j];
    advance(s0, sd_stage, b.ks + j, b.s + j, io.s_o + j);
    advance(sd0, sdd, b.ksd + j, b.sd + j, io.sd_o + j);
    return;
  }
  int r = k - n;
  if (r < 4) {
    const long long j = e * 4 + r;
    T q0;
    if (stage < 0) {
      const T* q = io.q + e * 4;
      const T nn = sqrt_t(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      q0 = q[r] / ((nn == T(0)) ? T(1) : nn);
      b.q0[j] = q0;
    } else {
      q0 = b.q0[j];
    }
    advance(q0, stage < 0 ? T(0) : b.qd[j], b.kq + j, b.q + j, io.q_o + j);
    return;
  }
  r -= 4;
  if (r < 9) {
    const int a = r / 3, x = r - 3 * a;
    const long long j = e * 3 + x;
    if (a == 0) advance(io.p[j], stage < 0 ? T(0) : b.pd[j], b.kp + j, b.p + j, io.p_o + j);
    else if (a == 1) advance(io.vl[j], stage < 0 ? T(0) : b.vd[e * 6 + x], b.kvl + j, b.vl + j, io.vl_o + j);
    else advance(io.om[j], stage < 0 ? T(0) : b.vd[e * 6 + 3 + x], b.kom + j, b.om + j, io.om_o + j);
    return;
  }
  r -= 9;
  {
    const long long j = e * 3 * nc + r;
    advance(io.m ? io.m[j] : T(0), stage < 0 ? T(0) : b.md[j], b.km + j, b.m + j, io.m_o ? io.m_o + j : b.m + j);
  }
}

template <typename T>
int dyn_t(const B200SimModel* m, int dtype, int64_t B, const T* s, const T* sd, const T* q, const T* vlin, const T* omega,
          const T* p, const T* mt, const T* tau, const T* fext, T* pd, T* qd, T* W_vd, T* sdd, T* md, void* stream) {
  Params<T> P;
  std::memset(&P, 0, sizeof(P));
  fill_model_params(m, P);
  P.B = B;
  P.s = s; P.sd = sd; P.q = q; P.vlin = vlin; P.omega = omega; P.p = p; P.m = mt; P.tau = tau; P.fext = fext;
  P.p_o = pd; P.q_o = qd; P.avd = W_vd; P.sdd_o = sdd; P.m_o = md;
  P.nsteps = 1; P.mode = MODE_DYN;
  return launch(m, P, dtype, stream);
}

template <typename T>
int step_rk4_t(B200SimModel* m, int dtype, int64_t B, Rk4Io<T> io, const T* fext, T* W_H_B, T* iXl, T* W_H_L, T* W_v, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int n = m->n, nc = (m->contact_model == 1) ? m->nc : 0;  // the contact state exists for SoftContacts only
  const size_t per_env = (size_t)(2 * n + 13 + 3 * nc)      // stage state
                         + (size_t)(3 + 4 + 6 + n + 3 * nc)  // slopes
                         + (size_t)(2 * n + 13 + 3 * nc)     // sums
                         + 4 + (size_t)n;                    // q0, tau
  const size_t bytes = (per_env * (size_t)B + 64) * sizeof(T) + 21 * 16;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev != m->device) CK(cudaSetDevice(m->device));
  void* base = nullptr;
  {
    std::lock_guard<std::mutex> lock(m->rigid_mutex);
    B200SimModel::StageScratch& sc = m->rk4_scratch[(void*)st];
    if (!sc.buf || sc.bytes < bytes) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone)
        return B200SIM_E_UNSUPPORTED;  // cannot allocate inside a capture: run one eager step first
      if (sc.buf) {
        CK(cudaStreamSynchronize(st));
        CK(cudaFree(sc.buf));
      }
      sc.buf = nullptr;
      sc.bytes = 0;
      CK(cudaMalloc(&sc.buf, bytes));
      sc.bytes = bytes;
    }
    base = sc.buf;
  }
  char* cur = (char*)base;
  auto take = [&](size_t count) {
    T* ptr = (T*)cur;
    cur += ((count * sizeof(T) + 15) / 16) * 16;  // every leaf 16-byte aligned
    return ptr;
  };
  Rk4Bufs<T> b;
  const size_t Bn = (size_t)B * n, Bm = (size_t)B * 3 * nc, B3 = (size_t)B * 3, B4 = (size_t)B * 4;
  b.s = take(Bn); b.sd = take(Bn); b.q = take(B4); b.vl = take(B3); b.om = take(B3); b.p = take(B3); b.m = take(Bm);
  b.pd = take(B3); b.qd = take(B4); b.vd = take((size_t)B * 6); b.sdd = take(Bn); b.md = take(Bm);
  b.ks = take(Bn); b.ksd = take(Bn); b.kq = take(B4); b.kvl = take(B3); b.kom = take(B3); b.kp = take(B3); b.km = take(Bm);
  b.q0 = take(B4); b.tau = take(Bn);
  const int E = n + 13 + 3 * nc;
  const long long total = (long long)B * E;
  const int threads = 256;
  const unsigned grid = (unsigned)((total + threads - 1) / threads);
  const T dt = (T)m->dt;
  int rc = 0;
  for (int stage = -1; stage <= 3 && !rc; ++stage) {
    if (stage >= 0)
      rc = dyn_t<T>(m, dtype, B, b.s, b.sd, b.q, b.vl, b.om, b.p, nc ? b.m : nullptr, b.tau, fext, b.pd, b.qd, b.vd, b.sdd,
                    nc ? b.md : nullptr, stream);
    if (rc) break;
    rk4_stage_kernel<T><<<grid, threads, 0, st>>>(B, n, nc, stage, dt, io, b, Blob<T>::cst(m), m->enable_friction, (T)m->tau_max,
                                                   (T)m->w_th, (T)m->w_max);
    rc = (int)cudaGetLastError();
  }
  // data.replace (api/data.py:441-447): quaternion normalised, caches of the new state
  if (!rc) rc = fk_t<T>(m, dtype, B, io.s_o, io.sd_o, io.q_o, io.vl_o, io.om_o, io.p_o, io.q_o, W_H_B, iXl, W_H_L, W_v, stream);
  if (dev != m->device) cudaSetDevice(dev);
  return rc;
}

}  // namespace

// ---- forward-mode AD (Dual<double>) -------------------------------------------------
int upload_dual(const std::vector<double>& val, const std::vector<double>& tan, DualD** d, cudaStream_t st) {
  std::vector<DualD> tmp(((val.size() + 3) & ~size_t(3)) + 4, DualD(0.0, 0.0));
  for (size_t i = 0; i < val.size(); ++i) tmp[i] = DualD(val[i], tan.empty() ? 0.0 : tan[i]);
  if (!*d) CK(cudaMalloc((void**)d, tmp.size() * sizeof(DualD)));
  // pageable source: the runtime stages it before returning, the device-side copy is ordered on `st`
  CK(cudaMemcpyAsync(*d, tmp.data(), tmp.size() * sizeof(DualD), cudaMemcpyHostToDevice, st));
  return 0;
}

int launch_dual(const B200SimModel* m, Params<DualD>& P, void* stream) {
  int fG = 8;  // lanes fixed to the instantiated widths
  while (fG > 1 && fG / 2 >= m->nL) fG /= 2;
  Geometry g;
  int rc = pick_geometry(m, 2, P.B, &g, fG);
  if (rc) return rc;
  P.envs_per_block = g.epb;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev != m->device) CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  switch (g.G) {
    case 1: rc = launch_g<DualD, 1>(P, g, st); break;
    case 2: rc = launch_g<DualD, 2>(P, g, st); break;
    case 4: rc = launch_g<DualD, 4>(P, g, st); break;
    case 8: rc = launch_g<DualD, 8>(P, g, st); break;
    default: rc = B200SIM_E_INVALID;
  }
  if (dev != m->device) cudaSetDevice(dev);
  return rc;
}

// (value, tangent) image of the model constants for a mass direction dm (HOST, nL; NULL = no direction):
// d(mass) = dm, d(D_link) = dm (|c|^2 1 - c c^T) (Inertia.to_sixd with the CoM and the CoM inertia held fixed); everything
// else is constant.  The image is rewritten only when its tangent part changes: a direction is given, or the resident
// image still carries the previous call's.  The copy is ordered on the caller's stream (JVPs of one model are expected
// on one stream), so consecutive directions need no host synchronisation.
int ensure_dual_blobs(B200SimModel* m, const double* link_mass_tangent, void* stream) {
  int prev = 0;
  CK(cudaGetDevice(&prev));
  CK(cudaSetDevice(m->device));
  int rc = 0;
  if (!m->cst_dd || link_mass_tangent || m->cst_dd_has_tangent) {
    std::vector<double> tan(m->cst_h.size(), 0.0);
    if (link_mass_tangent) {
      for (int i = 0; i < m->nL; ++i) {
        const double* c = m->cst_h.data() + (size_t)i * CREC;
        double* tc = tan.data() + (size_t)i * CREC;
        const double dm = link_mass_tangent[i];
        const double cx = c[C_COM], cy = c[C_COM + 1], cz = c[C_COM + 2];
        const double cc = cx * cx + cy * cy + cz * cz;
        tc[C_MASS] = dm;
        tc[C_DL + 0] = dm * (cc - cx * cx); tc[C_DL + 1] = -dm * cx * cy; tc[C_DL + 2] = -dm * cx * cz;
        tc[C_DL + 3] = dm * (cc - cy * cy); tc[C_DL + 4] = -dm * cy * cz; tc[C_DL + 5] = dm * (cc - cz * cz);
      }
    }
    rc = upload_dual(m->cst_h, tan, &m->cst_dd, (cudaStream_t)stream);
    m->cst_dd_has_tangent = link_mass_tangent != nullptr;
  }
  if (!rc && !m->csuc_dd) rc = upload_dual(m->csuc_h, {}, &m->csuc_dd, (cudaStream_t)stream);
  if (!rc && !m->pt_dd) rc = upload_dual(m->pt_h, {}, &m->pt_dd, (cudaStream_t)stream);
  cudaSetDevice(prev);
  return rc;
}

// ---- b200sim_step_vjp: gradient of <cotangent, step(x, theta)> w.r.t. joint positions and link masses ----------------
// Forward-mode columns contracted on the device.  The joint directions run as replicas of the batch in ONE launch of the
// forward-mode step kernel (replica r of environment b carries the unit tangent of joint j0 + r); a mass direction
// changes the shared model constants and takes a launch of its own over the un-replicated batch.
struct VjpIn { const double *s, *sd, *q, *vl, *om, *p, *m, *tau; };
struct VjpCt { const double *s, *sd, *q, *vl, *om, *p, *m; };
struct VjpDual { DualD *s, *sd, *q, *vl, *om, *p, *m, *tau, *s_o, *sd_o, *q_o, *vl_o, *om_o, *p_o, *m_o; };

__global__ void vjp_pack_kernel(long long B, int K, int n, int nc, int j0, VjpIn in, VjpDual d) {
  const int E = 3 * n + 13 + 3 * nc;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (long long)K * B * E) return;
  const long long e = tid / E;  // dual environment = replica * B + environment
  int k = (int)(tid - e * E);
  const long long b = e % B;
  const int r = (int)(e / B);
  if (k < n) { d.s[e * n + k] = DualD(in.s[b * n + k], (j0 >= 0 && k == j0 + r) ? 1.0 : 0.0); return; }
  k -= n;
  if (k < n) { d.sd[e * n + k] = DualD(in.sd[b * n + k], 0.0); return; }
  k -= n;
  if (k < n) { d.tau[e * n + k] = DualD(in.tau ? in.tau[b * n + k] : 0.0, 0.0); return; }
  k -= n;
  if (k < 4) { d.q[e * 4 + k] = DualD(in.q[b * 4 + k], 0.0); return; }
  k -= 4;
  if (k < 3) { d.vl[e * 3 + k] = DualD(in.vl[b * 3 + k], 0.0); return; }
  k -= 3;
  if (k < 3) { d.om[e * 3 + k] = DualD(in.om[b * 3 + k], 0.0); return; }
  k -= 3;
  if (k < 3) { d.p[e * 3 + k] = DualD(in.p[b * 3 + k], 0.0); return; }
  k -= 3;
  d.m[e * 3 * nc + k] = DualD(in.m ? in.m[b * 3 * nc + k] : 0.0, 0.0);
}

// g[b * g_stride + g_col0 + r] = sum over the output leaves of cotangent[b, :] . tangent_out[replica r, b, :]
__global__ void vjp_contract_kernel(long long B, int K, int n, int nc, VjpCt ct, VjpDual d, double* g, int g_stride, int g_col0) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)K * B) return;
  const long long b = e % B;
  const int r = (int)(e / B);
  double acc = 0.0;
  if (ct.s) for (int k = 0; k < n; ++k) acc += ct.s[b * n + k] * d.s_o[e * n + k].d;
  if (ct.sd) for (int k = 0; k < n; ++k) acc += ct.sd[b * n + k] * d.sd_o[e * n + k].d;
  if (ct.q) for (int k = 0; k < 4; ++k) acc += ct.q[b * 4 + k] * d.q_o[e * 4 + k].d;
  if (ct.vl) for (int k = 0; k < 3; ++k) acc += ct.vl[b * 3 + k] * d.vl_o[e * 3 + k].d;
  if (ct.om) for (int k = 0; k < 3; ++k) acc += ct.om[b * 3 + k] * d.om_o[e * 3 + k].d;
  if (ct.p) for (int k = 0; k < 3; ++k) acc += ct.p[b * 3 + k] * d.p_o[e * 3 + k].d;
  if (ct.m) for (int k = 0; k < 3 * nc; ++k) acc += ct.m[b * 3 * nc + k] * d.m_o[e * 3 * nc + k].d;
  g[b * g_stride + g_col0 + r] = acc;
}

int step_vjp_impl(B200SimModel* m, int64_t B, const VjpIn& in, const VjpCt& ct, double* g_s, double* g_mass, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int n = m->n, nL = m->nL, nc = (m->contact_model == 1) ? m->nc : 0;
  // replicas per launch: enough dual environments to keep the forward-mode kernel in its throughput regime
  const int Kc = std::max(1, std::min(std::max(n, nL), (int)(131072 / std::max<int64_t>(B, 1))));
  const size_t envs = (size_t)Kc * (size_t)B;
  const size_t per_env = (size_t)(3 * n + 13 + 3 * nc) + (size_t)(2 * n + 13 + 3 * nc);
  const size_t bytes = per_env * envs * sizeof(DualD) + 16 * 16;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev != m->device) CK(cudaSetDevice(m->device));
  void* base = nullptr;
  {
    std::lock_guard<std::mutex> lock(m->rigid_mutex);
    B200SimModel::StageScratch& sc = m->vjp_scratch[(void*)st];
    if (!sc.buf || sc.bytes < bytes) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) return B200SIM_E_UNSUPPORTED;
      if (sc.buf) {
        CK(cudaStreamSynchronize(st));
        CK(cudaFree(sc.buf));
      }
      sc.buf = nullptr;
      sc.bytes = 0;
      CK(cudaMalloc(&sc.buf, bytes));
      sc.bytes = bytes;
    }
    base = sc.buf;
  }
  DualD* cur = (DualD*)base;
  auto take = [&](size_t count) { DualD* ptr = cur; cur += count; return ptr; };  // sizeof(DualD) = 16: every leaf aligned
  VjpDual d;
  d.s = take(envs * n); d.sd = take(envs * n); d.q = take(envs * 4); d.vl = take(envs * 3); d.om = take(envs * 3);
  d.p = take(envs * 3); d.m = take(envs * 3 * nc); d.tau = take(envs * n);
  d.s_o = take(envs * n); d.sd_o = take(envs * n); d.q_o = take(envs * 4); d.vl_o = take(envs * 3); d.om_o = take(envs * 3);
  d.p_o = take(envs * 3); d.m_o = take(envs * 3 * nc);
  const int E = 3 * n + 13 + 3 * nc;
  const int threads = 256;
  auto launch_step = [&](long long Bd, long long mass_period = 0, int mass_first = 0) {
    Params<DualD> P;
    std::memset(&P, 0, sizeof(P));
    fill_model_params(m, P);
    P.flags &= ~F_TMA_STORE;
    P.B = Bd;
    P.mass_dir_period = mass_period;
    P.mass_dir_first = mass_first;
    P.s = d.s; P.sd = d.sd; P.q = d.q; P.vlin = d.vl; P.omega = d.om; P.p = d.p; P.m = nc ? d.m : nullptr; P.tau = d.tau;
    P.s_o = d.s_o; P.sd_o = d.sd_o; P.q_o = d.q_o; P.vlin_o = d.vl_o; P.omega_o = d.om_o; P.p_o = d.p_o;
    P.m_o = nc ? d.m_o : nullptr;
    P.nsteps = 1;
    P.mode = MODE_STEP;
    return launch_dual(m, P, stream);
  };
  int rc = 0;
  if (g_s && n > 0) {
    rc = ensure_dual_blobs(m, nullptr, stream);
    for (int j0 = 0; j0 < n && !rc; j0 += Kc) {
      const int K = std::min(Kc, n - j0);
      const long long total = (long long)K * B * E;
      vjp_pack_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(B, K, n, nc, j0, in, d);
      rc = (int)cudaGetLastError();
      if (!rc) rc = launch_step((long long)K * B);
      if (rc) break;
      vjp_contract_kernel<<<(unsigned)(((long long)K * B + threads - 1) / threads), threads, 0, st>>>(B, K, n, nc, ct, d, g_s, n, j0);
      rc = (int)cudaGetLastError();
    }
  }
  if (!rc && g_mass) {
    // the tangent part of the constants holds the unit mass direction of EVERY link; replica r keeps it for link k0 + r
    std::vector<double> ones(nL, 1.0);
    rc = ensure_dual_blobs(m, ones.data(), stream);
    for (int k0 = 0; k0 < nL && !rc; k0 += Kc) {
      const int K = std::min(Kc, nL - k0);
      const long long total = (long long)K * B * E;
      vjp_pack_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(B, K, n, nc, -1, in, d);
      rc = (int)cudaGetLastError();
      if (!rc) rc = launch_step((long long)K * B, B, k0);
      if (rc) break;
      vjp_contract_kernel<<<(unsigned)(((long long)K * B + threads - 1) / threads), threads, 0, st>>>(B, K, n, nc, ct, d, g_mass, nL, k0);
      rc = (int)cudaGetLastError();
    }
  }
  if (dev != m->device) cudaSetDevice(dev);
  return rc;
}

// Per-environment status flags of a step (b200sim_step_n_status): the conditions the reference can only raise as
// exceptions under JAXSIM_ENABLE_EXCEPTIONS (rbda/utils.py:136-146), evaluated on the device after the step.
template <typename T>
__global__ void status_kernel(long long B, int n, const T* __restrict__ q_in, const T* __restrict__ s_o, const T* __restrict__ sd_o,
                              const T* __restrict__ q_o, const T* __restrict__ vl_o, const T* __restrict__ om_o,
                              const T* __restrict__ p_o, int* __restrict__ status) {
  const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= B) return;
  int f = 0;
  const T* q = q_in + env * 4;
  const T qq = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (!(qq == qq)) f |= B200SIM_STATUS_QUATERNION_NAN;
  else if (!(fabs((double)qq - 1.0) <= 1e-8 + 1e-5)) f |= B200SIM_STATUS_QUATERNION_NOT_UNIT;  // jnp.allclose(q.q, 1.0)
  T acc = T(0);  // x - x is 0 for finite x, NaN for +-Inf and NaN
  for (int j = 0; j < n; ++j) { const T a = s_o[env * n + j], b = sd_o[env * n + j]; acc += (a - a) + (b - b); }
  for (int j = 0; j < 4; ++j) { const T a = q_o[env * 4 + j]; acc += a - a; }
  for (int j = 0; j < 3; ++j) { const T a = vl_o[env * 3 + j], b = om_o[env * 3 + j], c = p_o[env * 3 + j]; acc += (a - a) + (b - b) + (c - c); }
  if (!(acc == T(0))) f |= B200SIM_STATUS_NON_FINITE;
  if (f) status[env] |= f;
}

extern "C" {

const char* b200sim_version(void) { return "b200sim abi1 sm_100a fused-step"; }

int b200sim_model_create(const B200SimModelDesc* d, int device, B200SimModel** out) {
  if (!d || !out) return B200SIM_E_INVALID;
  if (d->abi_version != B200SIM_ABI_VERSION) return B200SIM_E_INVALID;
  const int nL = d->n_links, n = d->n_dofs, nc = d->n_points;
  if (nL < 1 || n != nL - 1 || nc < 0) return B200SIM_E_INVALID;
  if (!d->parent || !d->joint_type || !d->lam_H_pre || !d->suc_H_i || !d->joint_axis || !d->link_mass ||
      !d->link_com || !d->link_inertia)
    return B200SIM_E_INVALID;
  if (n > 0 && (!d->friction_static || !d->friction_viscous || !d->position_limits_min || !d->position_limits_max ||
                !d->position_limit_spring || !d->position_limit_damper))
    return B200SIM_E_INVALID;
  if (nc > 0 && (!d->point_body || !d->point_position || !d->point_enabled)) return B200SIM_E_INVALID;
  if (d->contact_model != B200SIM_CONTACT_NONE && d->contact_model != B200SIM_CONTACT_SOFT &&
      d->contact_model != B200SIM_CONTACT_RIGID && d->contact_model != B200SIM_CONTACT_RELAXED_RIGID)
    return B200SIM_E_UNSUPPORTED;
  if (d->contact_model >= B200SIM_CONTACT_RIGID && nc > 0) {
    // rigid.py:401-409 indexes the enabled subset twice: only a prefix makes that the identity
    bool seen_disabled = false;
    for (int k = 0; k < nc; ++k) {
      if (!d->point_enabled[k]) seen_disabled = true;
      else if (seen_disabled) return B200SIM_E_UNSUPPORTED;
    }
    if (!d->floating_base || !is_identity4(d->suc_H_i)) return B200SIM_E_UNSUPPORTED;
  }
  if (d->parent[0] != -1) return B200SIM_E_INVALID;
  for (int i = 1; i < nL; ++i) {
    if (d->parent[i] < 0 || d->parent[i] >= i) return B200SIM_E_INVALID;  // lambda(i) < i
    if (d->joint_type[i] != 1 && d->joint_type[i] != 2) return B200SIM_E_UNSUPPORTED;
  }
  for (int k = 0; k < nc; ++k)
    if (d->point_body[k] < 0 || d->point_body[k] >= nL) return B200SIM_E_INVALID;

  B200SimModel* m = new (std::nothrow) B200SimModel();
  if (!m) return B200SIM_E_INVALID;
  m->device = device;
  m->nL = nL; m->n = n; m->nc = nc;
  m->floating = d->floating_base ? 1 : 0;
  m->contact_model = d->contact_model;
  m->enable_friction = d->enable_friction ? 1 : 0;
  m->dt = d->time_step; m->g = d->gravity; m->h_terrain = d->terrain_height;
  m->K = d->soft_K; m->D = d->soft_D; m->mu = d->soft_mu; m->pexp = d->soft_p; m->qexp = d->soft_q;
  m->tau_max = d->torque_max; m->w_th = d->omega_th; m->w_max = d->omega_max;
  m->reg = d->rigid_regularization;
  m->rx_tc = d->relaxed_time_constant; m->rx_zeta = d->relaxed_damping_coefficient; m->rx_dmin = d->relaxed_d_min;
  m->rx_dmax = d->relaxed_d_max; m->rx_width = d->relaxed_width; m->rx_mid = d->relaxed_midpoint; m->rx_pow = d->relaxed_power;

  // ---- per-link constants
  m->cst_h.assign((size_t)nL * CREC, 0.0);
  m->csuc_h.assign((size_t)nL * 12, 0.0);
  bool suc_nonid = false;
  for (int i = 0; i < nL; ++i) {
    double* c = m->cst_h.data() + (size_t)i * CREC;
    const double* H = d->lam_H_pre + 16 * (size_t)i;
    const double* Hs = d->suc_H_i + 16 * (size_t)i;
    double Rpre[9], ax[3];
    for (int r = 0; r < 3; ++r) {
      for (int cc = 0; cc < 3; ++cc) {
        Rpre[3 * r + cc] = H[4 * r + cc];
        m->csuc_h[(size_t)i * 12 + 3 * r + cc] = Hs[4 * r + cc];
      }
      c[C_TPRE + r] = H[4 * r + 3];
      m->csuc_h[(size_t)i * 12 + 9 + r] = Hs[4 * r + 3];
      ax[r] = d->joint_axis[3 * (size_t)i + r];
      c[C_AXIS + r] = ax[r];
    }
    // R_rel(s) = Rpre (cos I + sin S(a) + (1 - cos) a a^T) = M0 + cos M1 + sin M2
    const double Sa[9] = {0, -ax[2], ax[1], ax[2], 0, -ax[0], -ax[1], ax[0], 0};
    for (int r = 0; r < 3; ++r) {
      for (int cc = 0; cc < 3; ++cc) {
        double raa = 0, rsa = 0;
        for (int k = 0; k < 3; ++k) { raa += Rpre[3 * r + k] * ax[k] * ax[cc]; rsa += Rpre[3 * r + k] * Sa[3 * k + cc]; }
        if (i == 0) {
          // link 0 has no joint: its slots carry suc_H_i[0] (root -> base link pose)
          c[C_M0 + 3 * r + cc] = Hs[4 * r + cc];
          c[C_M1 + 3 * r + cc] = 0.0;
          c[C_M2 + 3 * r + cc] = 0.0;
        } else if (d->joint_type[i] == 1) {
          c[C_M0 + 3 * r + cc] = raa;
          c[C_M1 + 3 * r + cc] = Rpre[3 * r + cc] - raa;
          c[C_M2 + 3 * r + cc] = rsa;
        } else {
          c[C_M0 + 3 * r + cc] = Rpre[3 * r + cc];
          c[C_M1 + 3 * r + cc] = 0.0;
          c[C_M2 + 3 * r + cc] = 0.0;
        }
      }
      double ra = 0;
      for (int k = 0; k < 3; ++k) ra += Rpre[3 * r + k] * ax[k];
      c[C_RA + r] = (i >= 1 && d->joint_type[i] == 2) ? ra : 0.0;
      c[C_PAX + r] = (i >= 1) ? ra : 0.0;  // a rotation about a leaves a invariant
      if (i == 0) c[C_TPRE + r] = Hs[4 * r + 3];
    }
    if (i >= 1 && !is_identity4(Hs)) suc_nonid = true;
    if (i >= 1) {
      const int j = i - 1;
      c[C_KS] = d->position_limit_spring[j];
      c[C_KD] = d->position_limit_damper[j];
      c[C_SMIN] = d->position_limits_min[j];
      c[C_SMAX] = d->position_limits_max[j];
      c[C_KC] = d->friction_static[j];
      c[C_KV] = d->friction_viscous[j];
    }
  }
  fill_link_consts(m, d->link_mass, d->link_com, d->link_inertia);
  m->pt_h.assign(d->point_position, d->point_position + (size_t)nc * 3);

  m->flags = 0;
  if (suc_nonid) m->flags |= F_SUC_NONID;
  if (!m->floating || !is_identity4(d->suc_H_i)) m->flags |= F_GENERIC_FK;
  if (d->soft_p == 0.5) m->flags |= F_SQRT_P;
  if (d->soft_q == 0.5) m->flags |= F_SQRT_Q;

  // ---- integer tables
  std::vector<int> depth(nL, 0);
  int maxd = 0;
  for (int i = 1; i < nL; ++i) { depth[i] = depth[d->parent[i]] + 1; maxd = std::max(maxd, depth[i]); }
  m->depth = maxd;
  std::vector<int> lvl_start(maxd + 2, 0), lvl_links;
  for (int l = 0; l <= maxd; ++l) {
    lvl_start[l] = (int)lvl_links.size();
    for (int i = 0; i < nL; ++i) if (depth[i] == l) lvl_links.push_back(i);
  }
  lvl_start[maxd + 1] = (int)lvl_links.size();
  std::vector<int> child_start(nL + 1, 0), child_idx;
  for (int i = 0; i < nL; ++i) {
    child_start[i] = (int)child_idx.size();
    for (int c = 1; c < nL; ++c) if (d->parent[c] == i) child_idx.push_back(c);
  }
  child_start[nL] = (int)child_idx.size();
  std::vector<int> pt_start(nL + 1, 0), pt_idx;
  for (int i = 0; i < nL; ++i) {
    pt_start[i] = (int)pt_idx.size();
    for (int k = 0; k < nc; ++k) if (d->point_body[k] == i) pt_idx.push_back(k);
  }
  pt_start[nL] = (int)pt_idx.size();

  auto push = [&](const int* src, size_t cnt) {
    int off = (int)m->itab_h.size();
    m->itab_h.insert(m->itab_h.end(), src, src + cnt);
    return off;
  };
  std::vector<int> lvl_start2(lvl_start);  // lvl_start has depth+2 entries: [0..depth+1]
  m->o_parent = push(d->parent, nL);
  m->o_jtype = push(d->joint_type, nL);
  m->o_lvl_start = push(lvl_start2.data(), lvl_start2.size());
  m->o_lvl_links = push(lvl_links.data(), lvl_links.size());
  m->o_child_start = push(child_start.data(), child_start.size());
  m->o_child_idx = push(child_idx.data(), child_idx.size());
  m->o_pt_start = push(pt_start.data(), pt_start.size());
  m->o_pt_idx = push(pt_idx.data(), pt_idx.size());
  m->o_pt_body = push(d->point_body, nc);
  m->o_pt_enabled = push(d->point_enabled, nc);
  if (d->contact_model >= B200SIM_CONTACT_RIGID && nc > 0) {
    // ancestors of every link, root's child first, the link itself last
    std::vector<int> anc((size_t)nL * std::max(maxd, 1), 0);
    for (int i = 1; i < nL; ++i) {
      int j = i;
      for (int t = depth[i] - 1; t >= 0; --t) { anc[(size_t)i * maxd + t] = j; j = d->parent[j]; }
    }
    m->o_anc = push(anc.data(), anc.size());
    m->o_ldepth = push(depth.data(), depth.size());
  }
  // Packed rows of the level walks for the specialised step kernel: row = G consecutive links of
  // one tree level (levels in order, a level wider than G spans several rows), one int per lane:
  //   bits 0-7 link (0xFF: none) | 8-15 parent | 16-23 first child | 24-26 number of children | 27-28 joint type
  // One independent shared-memory load per row replaces the dependent chain lvl_start -> lvl_links ->
  // parent / child_start -> child_idx / jtype.  Needs nL <= 255 and contiguous child indices (any
  // breadth-first numbering, e.g. the reference's, parsers/kinematic_graph.py:669-709).
  {
    bool ok = nL <= 255;
    for (int i = 0; i < nL && ok; ++i) {
      const int c0 = child_start[i], c1 = child_start[i + 1];
      if (c1 - c0 > 7) ok = false;
      for (int c = c0; c + 1 < c1; ++c) if (child_idx[c + 1] != child_idx[c] + 1) ok = false;
    }
    for (int pass = 0; pass < 2 && ok; ++pass) {
      const int Gr = pass == 0 ? 8 : 16;
      std::vector<int> rows;
      for (int l = 1; l <= maxd; ++l) {
        const int b = lvl_start[l], e2 = lvl_start[l + 1];
        for (int r0 = b; r0 < e2; r0 += Gr) {
          for (int k = 0; k < Gr; ++k) {
            int ent = 0xFF;
            if (r0 + k < e2) {
              const int i = lvl_links[r0 + k];
              const int nch = child_start[i + 1] - child_start[i];
              const int fc = nch > 0 ? child_idx[child_start[i]] : 0;
              ent = i | (d->parent[i] << 8) | (fc << 16) | (nch << 24) | ((d->joint_type[i] & 3) << 27);
            }
            rows.push_back(ent);
          }
        }
      }
      const int nrows = (int)rows.size() / Gr;
      const int off = rows.empty() ? 0 : push(rows.data(), rows.size());
      if (pass == 0) { m->o_rows8 = off; m->n_rows8 = nrows; } else { m->o_rows16 = off; m->n_rows16 = nrows; }
      // step2_kernel format: the children ADD their contribution into the parent's record, siblings that share a
      // row take turns in the order of their rank:
      //   bits 0-7 link | 8-15 parent | 16-19 rank among the siblings of this row | 20-23 sub-rounds of the row | 27-28 joint type
      std::vector<int> rows2(rows.size(), 0xFF);
      for (int r = 0; r < nrows; ++r) {
        int nsub = 1;
        std::vector<int> rank(Gr, 0);
        for (int k = 0; k < Gr; ++k) {
          const int ent = rows[(size_t)r * Gr + k];
          if ((ent & 0xFF) == 0xFF) continue;
          const int par = (ent >> 8) & 0xFF;
          for (int k2 = 0; k2 < k; ++k2) {
            const int e2 = rows[(size_t)r * Gr + k2];
            if ((e2 & 0xFF) != 0xFF && ((e2 >> 8) & 0xFF) == par) ++rank[k];
          }
          nsub = std::max(nsub, rank[k] + 1);
        }
        if (nsub > 15) ok = false;
        for (int k = 0; k < Gr; ++k) {
          const int ent = rows[(size_t)r * Gr + k];
          int e2 = 0xFF | (nsub << 20);
          if ((ent & 0xFF) != 0xFF) e2 = (ent & 0xFFFF) | (rank[k] << 16) | (nsub << 20) | (ent & (3 << 27));
          rows2[(size_t)r * Gr + k] = e2;
        }
      }
      const int off2 = rows2.empty() ? 0 : push(rows2.data(), rows2.size());
      if (pass == 0) m->o_rows2_8 = off2; else m->o_rows2_16 = off2;
    }
    if (!ok) { m->n_rows8 = 0; m->n_rows16 = 0; }
  }
  if (m->n_rows8 > 0 && m->n_rows16 > 0) {
    auto push2 = [&](const int* src, size_t cnt) {
      int off = (int)m->itab2_h.size();
      m->itab2_h.insert(m->itab2_h.end(), src, src + cnt);
      return off;
    };
    const int* it = m->itab_h.data();
    m->t2_parent = push2(it + m->o_parent, nL);
    m->t2_jtype = push2(it + m->o_jtype, nL);
    m->t2_pt_start = push2(it + m->o_pt_start, nL + 1);
    m->t2_pt_idx = push2(it + m->o_pt_idx, nc);
    m->t2_pt_body = push2(it + m->o_pt_body, nc);
    m->t2_pt_enabled = push2(it + m->o_pt_enabled, nc);
    m->t2_rows8 = push2(it + m->o_rows2_8, (size_t)m->n_rows8 * 8);
    m->t2_rows16 = push2(it + m->o_rows2_16, (size_t)m->n_rows16 * 16);
  }
  if (m->itab_h.empty()) m->itab_h.push_back(0);

  // ---- upload
  int prev = 0;
  cudaError_t e = cudaGetDevice(&prev);
  if (e != cudaSuccess) { delete m; return (int)e; }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) { delete m; return (int)e; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { delete m; return (int)e; }
  m->num_sms = prop.multiProcessorCount;
  m->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  int rc = 0;
  if (!rc) rc = upload(m->cst_h, &m->cst_f);
  if (!rc) rc = upload(m->cst_h, &m->cst_d);
  if (!rc) rc = upload(m->csuc_h, &m->csuc_f);
  if (!rc) rc = upload(m->csuc_h, &m->csuc_d);
  if (!rc) rc = upload(m->pt_h, &m->pt_f);
  if (!rc) rc = upload(m->pt_h, &m->pt_d);
  if (!rc) {
    std::vector<int> padded(((m->itab_h.size() + 3) & ~size_t(3)) + 4, 0);
    std::copy(m->itab_h.begin(), m->itab_h.end(), padded.begin());
    e = cudaMalloc((void**)&m->itab_d, padded.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(m->itab_d, padded.data(), padded.size() * sizeof(int), cudaMemcpyHostToDevice);
    rc = (int)e;
  }
  if (!rc && !m->itab2_h.empty()) {
    std::vector<int> padded(((m->itab2_h.size() + 3) & ~size_t(3)) + 4, 0);
    std::copy(m->itab2_h.begin(), m->itab2_h.end(), padded.begin());
    e = cudaMalloc((void**)&m->itab2_d, padded.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(m->itab2_d, padded.data(), padded.size() * sizeof(int), cudaMemcpyHostToDevice);
    rc = (int)e;
  }
  cudaSetDevice(prev);
  if (rc) { b200sim_model_destroy(m); return rc; }
  // the model must fit at least one environment per block in both precisions we may be asked for
  Geometry g;
  rc = pick_geometry(m, B200SIM_DTYPE_F32, 1, &g);
  if (rc) { b200sim_model_destroy(m); return rc; }
  *out = m;
  return 0;
}

void b200sim_model_destroy(B200SimModel* m) {
  if (!m) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(m->device);
  cudaFree(m->cst_f); cudaFree(m->cst_d); cudaFree(m->csuc_f); cudaFree(m->csuc_d);
  cudaFree(m->pt_f); cudaFree(m->pt_d); cudaFree(m->itab_d); cudaFree(m->itab2_d); cudaFree(m->dbg_d);
  for (auto& kv : m->rigid_scratch) {
    cudaFree(kv.second.buf);
    cudaFree(kv.second.qp);
    if (kv.second.aux) { cudaStreamDestroy(kv.second.aux); cudaEventDestroy(kv.second.ev_fork); cudaEventDestroy(kv.second.ev_join); }
  }
  for (auto& kv : m->rk4_scratch) cudaFree(kv.second.buf);
  for (auto& kv : m->vjp_scratch) cudaFree(kv.second.buf);
  cudaFree(m->cst_dd); cudaFree(m->csuc_dd); cudaFree(m->pt_dd);
  cudaSetDevice(prev);
  delete m;
}

int b200sim_model_update_link_params(B200SimModel* m, const double* mass, const double* com, const double* inertia6) {
  if (!m || !mass || !com || !inertia6) return B200SIM_E_INVALID;
  fill_link_consts(m, mass, com, inertia6);
  int prev = 0;
  CK(cudaGetDevice(&prev));
  CK(cudaSetDevice(m->device));
  CK(cudaDeviceSynchronize());
  int rc = upload(m->cst_h, &m->cst_f);
  if (!rc) rc = upload(m->cst_h, &m->cst_d);
  m->cst_dd_has_tangent = true;  // the forward-mode image is stale: the next JVP rewrites it
  cudaSetDevice(prev);
  return rc;
}

int b200sim_model_set_tuning(B200SimModel* m, int lanes_per_env, int envs_per_block) {
  if (!m) return B200SIM_E_INVALID;
  const int G = lanes_per_env;
  if (!(G == 0 || G == 1 || G == 2 || G == 4 || G == 8 || G == 16 || G == 32)) return B200SIM_E_INVALID;
  if (envs_per_block < 0) return B200SIM_E_INVALID;
  m->tune_G = G;
  m->tune_epb = envs_per_block;
  return 0;
}

// Undeclared diagnostic (not part of the ABI): enables, reads and resets the rigid-contact counters
// [0] QP iterations, [1] QPs, [2] max iterations, [3] active points (QP), [4] full items,
// [5] impact-only items, [6] impacts, [7] active points (impact).
constexpr int DBG_WORDS = 8 + 32 + 1024 + 512 + 24 + 512;  // ... + the contact-problem dump of B200SIM_RIGID_DEBUG builds at word 1600  // 8 counters + the phase clocks of step_kernel (B200SIM_PHASE_MARK) + per-block start/end ns
extern "C" int b200sim_debug_counters(B200SimModel* m, unsigned long long* out8) {
  if (!m) return B200SIM_E_INVALID;
  int prev = 0;
  CK(cudaGetDevice(&prev));
  CK(cudaSetDevice(m->device));
  if (!m->dbg_d) {
    CK(cudaMalloc((void**)&m->dbg_d, DBG_WORDS * sizeof(unsigned long long)));
    CK(cudaMemset(m->dbg_d, 0, DBG_WORDS * sizeof(unsigned long long)));
  }
  CK(cudaDeviceSynchronize());
  if (out8) CK(cudaMemcpy(out8, m->dbg_d, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  CK(cudaMemset(m->dbg_d, 0, 8 * sizeof(unsigned long long)));
  cudaSetDevice(prev);
  return 0;
}

// Undeclared diagnostic: the phase clocks (clock64 of one warp, see B200SIM_PHASE_MARK) of the last
// step_kernel launch; requires b200sim_debug_counters to have been called once (enables the buffer).
extern "C" int b200sim_debug_phase_clocks(B200SimModel* m, unsigned long long* out32) {
  if (!m || !m->dbg_d || !out32) return B200SIM_E_INVALID;
  int prev = 0;
  CK(cudaGetDevice(&prev));
  CK(cudaSetDevice(m->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out32, m->dbg_d + 8, 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  cudaSetDevice(prev);
  return 0;
}

// Undeclared diagnostic: %globaltimer (ns) at the start and end of the first 512 blocks of the last
// step_kernel launch: out[2*b], out[2*b+1].
extern "C" int b200sim_debug_block_times(B200SimModel* m, unsigned long long* out1024) {
  if (!m || !m->dbg_d || !out1024) return B200SIM_E_INVALID;
  int prev = 0;
  CK(cudaGetDevice(&prev));
  CK(cudaSetDevice(m->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out1024, m->dbg_d + 40, (1024 + 512) * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  cudaSetDevice(prev);
  return 0;
}

// Undeclared diagnostic (B200SIM_RIGID_DEBUG builds): the 512-double dump of environment 0's contact problem.
extern "C" int b200sim_debug_rigid_dump(B200SimModel* m, double* out512) {
  if (!m || !m->dbg_d || !out512) return B200SIM_E_INVALID;
  int prev = 0;
  CK(cudaGetDevice(&prev));
  CK(cudaSetDevice(m->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out512, m->dbg_d + 1600, 512 * sizeof(double), cudaMemcpyDeviceToHost));
  cudaSetDevice(prev);
  return 0;
}

int b200sim_model_set_options(B200SimModel* m, int32_t options) {
  if (!m || (options & ~(B200SIM_OPT_TMA_STORE | B200SIM_OPT_RIGID_QP_F32 | B200SIM_OPT_GENERIC_KERNEL | B200SIM_OPT_BULK_IN | B200SIM_OPT_NO_PDL | B200SIM_OPT_STEP_V1 | B200SIM_OPT_NO_BULK_IN | B200SIM_OPT_RIGID_MONO))) return B200SIM_E_INVALID;
  m->opt_flags = options;
  return 0;
}

int b200sim_model_query(const B200SimModel* m, int dtype, int64_t B, int32_t* G, int32_t* epb, int32_t* grid, int32_t* smem) {
  if (!m || B < 1 || (dtype != 0 && dtype != 1)) return B200SIM_E_INVALID;
  Geometry g;
  int rc = B200SIM_E_UNSUPPORTED;
  {
    // the geometry of the kernel a soft-contact / contact-free `step` of this model launches
    Params<float> Pq;
    std::memset(&Pq, 0, sizeof(Pq));
    Pq.mode = MODE_STEP; Pq.floating = m->floating; Pq.flags = m->flags; Pq.n = m->n; Pq.nc = m->nc; Pq.contact_model = m->contact_model;
    if (specialised_step_applies(m, Pq) && !(m->opt_flags & B200SIM_OPT_STEP_V1))
      rc = pick_geometry2(m, dtype == B200SIM_DTYPE_F64 ? 8 : 4, B, &g);
  }
  if (rc) rc = pick_geometry(m, dtype, B, &g);
  if (rc) return rc;
  if (G) *G = g.G;
  if (epb) *epb = g.epb;
  if (grid) *grid = g.grid;
  if (smem) *smem = (int32_t)g.smem;
  return 0;
}

static bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

int b200sim_step_n(const B200SimModel* m, int dtype, int64_t B, int32_t nsteps, const void* s, const void* sd,
                   const void* q, const void* vlin, const void* omega, const void* p, const void* mt, const void* tau,
                   int64_t tau_step_stride, const void* fext, int64_t fext_step_stride, const void* W_H_L_in,
                   const void* W_v_in, void* s_o, void* sd_o, void* q_o, void* vlin_o, void* omega_o, void* p_o,
                   void* m_o, void* W_H_B, void* iXl, void* W_H_L, void* W_v, void* stream) {
  if (!m || B < 0 || nsteps < 1 || (dtype != 0 && dtype != 1)) return B200SIM_E_INVALID;
  if (tau_step_stride < 0 || fext_step_stride < 0) return B200SIM_E_INVALID;
  if (B == 0) return 0;
  if (!q || !vlin || !omega || !p || !q_o || !vlin_o || !omega_o || !p_o) return B200SIM_E_INVALID;
  if (m->n > 0 && (!s || !sd || !s_o || !sd_o)) return B200SIM_E_INVALID;
  // vector / TMA stores: the cache outputs must be 16-byte aligned (8 for float W_v rows)
  if (!aligned(W_H_B, 16) || !aligned(iXl, 16) || !aligned(W_H_L, 16) || !aligned(W_v, dtype == 0 ? 8 : 16))
    return B200SIM_E_INVALID;
  if ((W_H_L_in == nullptr) != (W_v_in == nullptr)) return B200SIM_E_INVALID;
  if (!aligned(W_H_L_in, 16) || !aligned(W_v_in, dtype == 0 ? 8 : 16)) return B200SIM_E_INVALID;
  if (dtype == 0)
    return step_t<float>(m, dtype, B, s, sd, q, vlin, omega, p, mt, tau, fext, s_o, sd_o, q_o, vlin_o, omega_o, p_o, m_o,
                         W_H_B, iXl, W_H_L, W_v, nsteps, tau_step_stride, fext_step_stride, W_H_L_in, W_v_in, stream);
  return step_t<double>(m, dtype, B, s, sd, q, vlin, omega, p, mt, tau, fext, s_o, sd_o, q_o, vlin_o, omega_o, p_o, m_o,
                        W_H_B, iXl, W_H_L, W_v, nsteps, tau_step_stride, fext_step_stride, W_H_L_in, W_v_in, stream);
}

int b200sim_step_n_ex(const B200SimModel* m, int dtype, int64_t B, int32_t nsteps, const void* s, const void* sd,
                      const void* q, const void* vlin, const void* omega, const void* p, const void* mt, const void* tau,
                      int64_t tau_step_stride, const void* fext, int64_t fext_step_stride, const void* W_H_L_in,
                      const void* W_v_in, void* s_o, void* sd_o, void* q_o, void* vlin_o, void* omega_o, void* p_o,
                      void* m_o, void* W_H_B, void* iXl, void* W_H_L, void* W_v, int32_t f_ext_representation,
                      int32_t* status_flags, void* stream) {
  if (f_ext_representation < 0 || f_ext_representation > 2) return B200SIM_E_INVALID;
  // body-fixed / mixed forces are re-expressed with the ABA chain poses: available where they ARE the link poses
  if (m && f_ext_representation != 0 && fext && (m->flags & F_GENERIC_FK)) return B200SIM_E_UNSUPPORTED;
  if (!status_flags && (f_ext_representation == 0 || !fext))
    return b200sim_step_n(m, dtype, B, nsteps, s, sd, q, vlin, omega, p, mt, tau, tau_step_stride, fext, fext_step_stride,
                          W_H_L_in, W_v_in, s_o, sd_o, q_o, vlin_o, omega_o, p_o, m_o, W_H_B, iXl, W_H_L, W_v, stream);
  if (!m || B < 0 || nsteps < 1 || (dtype != 0 && dtype != 1)) return B200SIM_E_INVALID;
  if (tau_step_stride < 0 || fext_step_stride < 0) return B200SIM_E_INVALID;
  if (B == 0) return 0;
  if (!q || !vlin || !omega || !p || !q_o || !vlin_o || !omega_o || !p_o) return B200SIM_E_INVALID;
  if (m->n > 0 && (!s || !sd || !s_o || !sd_o)) return B200SIM_E_INVALID;
  if (!aligned(W_H_B, 16) || !aligned(iXl, 16) || !aligned(W_H_L, 16) || !aligned(W_v, dtype == 0 ? 8 : 16))
    return B200SIM_E_INVALID;
  if ((W_H_L_in == nullptr) != (W_v_in == nullptr)) return B200SIM_E_INVALID;
  if (!aligned(W_H_L_in, 16) || !aligned(W_v_in, dtype == 0 ? 8 : 16)) return B200SIM_E_INVALID;
  if (status_flags && q == q_o) return B200SIM_E_INVALID;  // the flags describe the INPUT quaternion too: no in-place step
  cudaStream_t st = (cudaStream_t)stream;
  int prev = 0;
  CK(cudaGetDevice(&prev));
  if (prev != m->device) CK(cudaSetDevice(m->device));
  int rc = status_flags ? (int)cudaMemsetAsync(status_flags, 0, (size_t)B * sizeof(int32_t), st) : 0;
  if (!rc) {
    rc = dtype == 0
             ? step_t<float>(m, dtype, B, s, sd, q, vlin, omega, p, mt, tau, fext, s_o, sd_o, q_o, vlin_o, omega_o, p_o, m_o, W_H_B,
                             iXl, W_H_L, W_v, nsteps, tau_step_stride, fext_step_stride, W_H_L_in, W_v_in, stream, status_flags,
                             f_ext_representation)
             : step_t<double>(m, dtype, B, s, sd, q, vlin, omega, p, mt, tau, fext, s_o, sd_o, q_o, vlin_o, omega_o, p_o, m_o, W_H_B,
                              iXl, W_H_L, W_v, nsteps, tau_step_stride, fext_step_stride, W_H_L_in, W_v_in, stream, status_flags,
                              f_ext_representation);
  }
  if (!rc && status_flags) {
    const int threads = 128;
    const int blocks = (int)((B + threads - 1) / threads);
    if (dtype == 0)
      status_kernel<float><<<blocks, threads, 0, st>>>(B, m->n, (const float*)q, (const float*)s_o, (const float*)sd_o, (const float*)q_o,
                                                       (const float*)vlin_o, (const float*)omega_o, (const float*)p_o, status_flags);
    else
      status_kernel<double><<<blocks, threads, 0, st>>>(B, m->n, (const double*)q, (const double*)s_o, (const double*)sd_o,
                                                        (const double*)q_o, (const double*)vlin_o, (const double*)omega_o,
                                                        (const double*)p_o, status_flags);
    rc = (int)cudaGetLastError();
  }
  if (prev != m->device) cudaSetDevice(prev);
  return rc;
}

int b200sim_step_n_status(const B200SimModel* m, int dtype, int64_t B, int32_t nsteps, const void* s, const void* sd,
                          const void* q, const void* vlin, const void* omega, const void* p, const void* mt, const void* tau,
                          int64_t tau_step_stride, const void* fext, int64_t fext_step_stride, const void* W_H_L_in,
                          const void* W_v_in, void* s_o, void* sd_o, void* q_o, void* vlin_o, void* omega_o, void* p_o,
                          void* m_o, void* W_H_B, void* iXl, void* W_H_L, void* W_v, int32_t* status_flags, void* stream) {
  return b200sim_step_n_ex(m, dtype, B, nsteps, s, sd, q, vlin, omega, p, mt, tau, tau_step_stride, fext, fext_step_stride,
                           W_H_L_in, W_v_in, s_o, sd_o, q_o, vlin_o, omega_o, p_o, m_o, W_H_B, iXl, W_H_L, W_v, 0, status_flags,
                           stream);
}

int b200sim_step(const B200SimModel* m, int dtype, int64_t B, const void* s, const void* sd, const void* q,
                 const void* vlin, const void* omega, const void* p, const void* mt, const void* tau, const void* fext,
                 void* s_o, void* sd_o, void* q_o, void* vlin_o, void* omega_o, void* p_o, void* m_o, void* W_H_B,
                 void* iXl, void* W_H_L, void* W_v, void* stream) {
  return b200sim_step_n(m, dtype, B, 1, s, sd, q, vlin, omega, p, mt, tau, 0, fext, 0, nullptr, nullptr, s_o, sd_o, q_o,
                        vlin_o, omega_o, p_o, m_o, W_H_B, iXl, W_H_L, W_v, stream);
}

int b200sim_fk(const B200SimModel* m, int dtype, int64_t B, const void* s, const void* sd, const void* q,
               const void* vlin, const void* omega, const void* p, void* q_o, void* W_H_B, void* iXl, void* W_H_L,
               void* W_v, void* stream) {
  if (!m || B < 0 || (dtype != 0 && dtype != 1)) return B200SIM_E_INVALID;
  if (B == 0) return 0;
  if (!q || !vlin || !omega || !p) return B200SIM_E_INVALID;
  if (m->n > 0 && (!s || !sd)) return B200SIM_E_INVALID;
  if (!aligned(W_H_B, 16) || !aligned(iXl, 16) || !aligned(W_H_L, 16) || !aligned(W_v, dtype == 0 ? 8 : 16))
    return B200SIM_E_INVALID;
  if (dtype == 0) return fk_t<float>(m, dtype, B, s, sd, q, vlin, omega, p, q_o, W_H_B, iXl, W_H_L, W_v, stream);
  return fk_t<double>(m, dtype, B, s, sd, q, vlin, omega, p, q_o, W_H_B, iXl, W_H_L, W_v, stream);
}

int b200sim_aba(const B200SimModel* m, int dtype, int64_t B, const void* s, const void* sd, const void* q,
                const void* vlin, const void* omega, const void* p, const void* tau, const void* fext, void* avd,
                void* sdd, void* stream) {
  if (!m || B < 0 || (dtype != 0 && dtype != 1)) return B200SIM_E_INVALID;
  if (B == 0) return 0;
  if (!q || !vlin || !omega || !p || !avd) return B200SIM_E_INVALID;
  if (m->n > 0 && (!s || !sd || !sdd)) return B200SIM_E_INVALID;
  if (dtype == 0) return aba_t<float>(m, dtype, B, s, sd, q, vlin, omega, p, tau, fext, avd, sdd, stream);
  return aba_t<double>(m, dtype, B, s, sd, q, vlin, omega, p, tau, fext, avd, sdd, stream);
}

int b200sim_rnea(const B200SimModel* m, int dtype, int64_t B, const void* s, const void* sd, const void* q,
                 const void* vlin, const void* omega, const void* p, const void* W_vd_WB, const void* sdd,
                 const void* fext, void* W_f_B, void* tau, void* stream) {
  if (!m || B < 0 || (dtype != 0 && dtype != 1)) return B200SIM_E_INVALID;
  if (B == 0) return 0;
  if (!q || !vlin || !omega || !p || !W_f_B) return B200SIM_E_INVALID;
  if (m->n > 0 && (!s || !sd || !tau)) return B200SIM_E_INVALID;
  if (dtype == 0) return rnea_t<float>(m, dtype, B, s, sd, q, vlin, omega, p, W_vd_WB, sdd, fext, W_f_B, tau, stream);
  return rnea_t<double>(m, dtype, B, s, sd, q, vlin, omega, p, W_vd_WB, sdd, fext, W_f_B, tau, stream);
}

int b200sim_crba(const B200SimModel* m, int dtype, int64_t B, const void* s, void* M, void* stream) {
  if (!m || B < 0 || (dtype != 0 && dtype != 1)) return B200SIM_E_INVALID;
  if (B == 0) return 0;
  if (!M || (m->n > 0 && !s)) return B200SIM_E_INVALID;
  if (dtype == 0) return crba_t<float>(m, dtype, B, s, M, stream);
  return crba_t<double>(m, dtype, B, s, M, stream);
}

int b200sim_step_jvp(B200SimModel* m, int64_t B, int32_t nsteps, const double* link_mass_tangent, const void* s,
                     const void* sd, const void* q, const void* vlin, const void* omega, const void* p, const void* mt,
                     const void* tau, void* s_o, void* sd_o, void* q_o, void* vlin_o, void* omega_o, void* p_o,
                     void* m_o, void* W_H_B, void* iXl, void* W_H_L, void* W_v, void* stream) {
  return b200sim_step_jvp_ex(m, B, nsteps, link_mass_tangent, 0, 0, s, sd, q, vlin, omega, p, mt, tau, s_o, sd_o, q_o, vlin_o,
                             omega_o, p_o, m_o, W_H_B, iXl, W_H_L, W_v, stream);
}

int b200sim_step_jvp_ex(B200SimModel* m, int64_t B, int32_t nsteps, const double* link_mass_tangent,
                        int64_t mass_direction_period, int32_t mass_direction_first_link, const void* s, const void* sd,
                        const void* q, const void* vlin, const void* omega, const void* p, const void* mt, const void* tau,
                        void* s_o, void* sd_o, void* q_o, void* vlin_o, void* omega_o, void* p_o, void* m_o, void* W_H_B,
                        void* iXl, void* W_H_L, void* W_v, void* stream) {
  if (m && m->contact_model >= B200SIM_CONTACT_RIGID && m->nc > 0) return B200SIM_E_UNSUPPORTED;
  if (!m || B < 0 || nsteps < 1 || mass_direction_period < 0) return B200SIM_E_INVALID;
  if (mass_direction_period > 0 && !link_mass_tangent) return B200SIM_E_INVALID;
  if (B == 0) return 0;
  if (!q || !vlin || !omega || !p || !q_o || !vlin_o || !omega_o || !p_o) return B200SIM_E_INVALID;
  if (m->n > 0 && (!s || !sd || !s_o || !sd_o)) return B200SIM_E_INVALID;
  int rc = ensure_dual_blobs(m, link_mass_tangent, stream);
  if (rc) return rc;
  Params<DualD> P;
  std::memset(&P, 0, sizeof(P));
  fill_model_params(m, P);
  P.flags &= ~F_TMA_STORE;
  P.B = B;
  typedef DualD T;
  P.s = (const T*)s; P.sd = (const T*)sd; P.q = (const T*)q; P.vlin = (const T*)vlin; P.omega = (const T*)omega;
  P.p = (const T*)p; P.m = (const T*)mt; P.tau = (const T*)tau;
  P.s_o = (T*)s_o; P.sd_o = (T*)sd_o; P.q_o = (T*)q_o; P.vlin_o = (T*)vlin_o; P.omega_o = (T*)omega_o;
  P.p_o = (T*)p_o; P.m_o = (T*)m_o;
  P.W_H_B = (T*)W_H_B; P.iXl = (T*)iXl; P.W_H_L = (T*)W_H_L; P.W_v = (T*)W_v;
  P.nsteps = nsteps;
  P.mode = MODE_STEP;
  P.mass_dir_period = mass_direction_period;
  P.mass_dir_first = mass_direction_first_link;
  return launch_dual(m, P, stream);
}

int b200sim_step_vjp(B200SimModel* m, int64_t B, const void* s, const void* sd, const void* q, const void* vlin,
                     const void* omega, const void* p, const void* mt, const void* tau, const void* ct_s, const void* ct_sd,
                     const void* ct_q, const void* ct_vlin, const void* ct_omega, const void* ct_p, const void* ct_m,
                     void* grad_s, void* grad_link_mass, void* stream) {
  if (m && m->contact_model >= B200SIM_CONTACT_RIGID && m->nc > 0) return B200SIM_E_UNSUPPORTED;
  if (!m || B < 0) return B200SIM_E_INVALID;
  if (B == 0) return 0;
  if (!q || !vlin || !omega || !p) return B200SIM_E_INVALID;
  if (m->n > 0 && (!s || !sd)) return B200SIM_E_INVALID;
  typedef const double* cd;
  VjpIn in = {(cd)s, (cd)sd, (cd)q, (cd)vlin, (cd)omega, (cd)p, (cd)mt, (cd)tau};
  VjpCt ct = {(cd)ct_s, (cd)ct_sd, (cd)ct_q, (cd)ct_vlin, (cd)ct_omega, (cd)ct_p, (m->contact_model == 1 && m->nc > 0) ? (cd)ct_m : nullptr};
  return step_vjp_impl(m, B, in, ct, (double*)grad_s, (double*)grad_link_mass, stream);
}

int b200sim_step_rk4(B200SimModel* m, int dtype, int64_t B, const void* s, const void* sd, const void* q, const void* vlin,
                     const void* omega, const void* p, const void* mt, const void* tau_ref, const void* fext, void* s_o,
                     void* sd_o, void* q_o, void* vlin_o, void* omega_o, void* p_o, void* m_o, void* W_H_B, void* iXl,
                     void* W_H_L, void* W_v, void* stream) {
  if (!m || B < 0 || (dtype != 0 && dtype != 1)) return B200SIM_E_INVALID;
  if (m->contact_model >= B200SIM_CONTACT_RIGID && m->nc > 0) return B200SIM_E_UNSUPPORTED;
  if (B == 0) return 0;
  if (!q || !vlin || !omega || !p || !q_o || !vlin_o || !omega_o || !p_o) return B200SIM_E_INVALID;
  if (m->n > 0 && (!s || !sd || !s_o || !sd_o)) return B200SIM_E_INVALID;
  if (!aligned(W_H_B, 16) || !aligned(iXl, 16) || !aligned(W_H_L, 16) || !aligned(W_v, dtype == 0 ? 8 : 16))
    return B200SIM_E_INVALID;
  if (dtype == 0) {
    typedef float T;
    Rk4Io<T> io = {(const T*)s, (const T*)sd, (const T*)q, (const T*)vlin, (const T*)omega, (const T*)p, (const T*)mt,
                   (const T*)tau_ref, (T*)s_o, (T*)sd_o, (T*)q_o, (T*)vlin_o, (T*)omega_o, (T*)p_o, (T*)m_o};
    return step_rk4_t<T>(m, dtype, B, io, (const T*)fext, (T*)W_H_B, (T*)iXl, (T*)W_H_L, (T*)W_v, stream);
  }
  typedef double T;
  Rk4Io<T> io = {(const T*)s, (const T*)sd, (const T*)q, (const T*)vlin, (const T*)omega, (const T*)p, (const T*)mt,
                 (const T*)tau_ref, (T*)s_o, (T*)sd_o, (T*)q_o, (T*)vlin_o, (T*)omega_o, (T*)p_o, (T*)m_o};
  return step_rk4_t<T>(m, dtype, B, io, (const T*)fext, (T*)W_H_B, (T*)iXl, (T*)W_H_L, (T*)W_v, stream);
}

int b200sim_dynamics(const B200SimModel* m, int dtype, int64_t B, const void* s, const void* sd, const void* q,
                     const void* vlin, const void* omega, const void* p, const void* mt, const void* tau,
                     const void* fext, void* pd, void* qd, void* W_vd, void* sdd, void* md, void* stream) {
  if (!m || B < 0 || (dtype != 0 && dtype != 1)) return B200SIM_E_INVALID;
  if (m->contact_model >= B200SIM_CONTACT_RIGID && m->nc > 0) return B200SIM_E_UNSUPPORTED;
  if (B == 0) return 0;
  if (!q || !vlin || !omega || !p || !W_vd) return B200SIM_E_INVALID;
  if (m->n > 0 && (!s || !sd || !sdd)) return B200SIM_E_INVALID;
  if (dtype == 0) {
    Params<float> P;
    std::memset(&P, 0, sizeof(P));
    fill_model_params(m, P);
    typedef float T;
    P.B = B;
    P.s = (const T*)s; P.sd = (const T*)sd; P.q = (const T*)q; P.vlin = (const T*)vlin; P.omega = (const T*)omega;
    P.p = (const T*)p; P.m = (const T*)mt; P.tau = (const T*)tau; P.fext = (const T*)fext;
    P.p_o = (T*)pd; P.q_o = (T*)qd; P.avd = (T*)W_vd; P.sdd_o = (T*)sdd; P.m_o = (T*)md;
    P.nsteps = 1; P.mode = MODE_DYN;
    return launch(m, P, dtype, stream);
  }
  Params<double> P;
  std::memset(&P, 0, sizeof(P));
  fill_model_params(m, P);
  typedef double T;
  P.B = B;
  P.s = (const T*)s; P.sd = (const T*)sd; P.q = (const T*)q; P.vlin = (const T*)vlin; P.omega = (const T*)omega;
  P.p = (const T*)p; P.m = (const T*)mt; P.tau = (const T*)tau; P.fext = (const T*)fext;
  P.p_o = (T*)pd; P.q_o = (T*)qd; P.avd = (T*)W_vd; P.sdd_o = (T*)sdd; P.m_o = (T*)md;
  P.nsteps = 1; P.mode = MODE_DYN;
  return launch(m, P, dtype, stream);
}

}  // extern "C"
