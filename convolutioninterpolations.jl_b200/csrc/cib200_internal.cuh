// cib200_internal.cuh - shared declarations of the sm_100a library (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>

#include "cib200.h"

#define CIB_N_PRE 101
#define CIB_TAB_ROWS 101

namespace cib {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;

#define CIB_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      cib::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CIB200_ECUDA;                                                                      \
    }                                                                                           \
  } while (0)

#define CIB_LAUNCHED() (cib::g_launches.fetch_add(1, std::memory_order_relaxed))

// Tuning / experiment switches read from the environment exist only in builds made with -DCIB200_EXPERIMENTS
// (python build.py --experiments).  The shipped library ignores the environment: no variable can change which
// kernel runs, or skip work.
#ifdef CIB200_EXPERIMENTS
#include <stdlib.h>
inline const char *tune_env(const char *name) { return getenv(name); }
#else
inline const char *tune_env(const char *) { return nullptr; }
#endif

// ---- kernel registry ------------------------------------------------------------------------
int kernel_eqs(int kernel);              // degree_to_equations.jl:13-25
int kernel_poly_nodes(int kernel);       // interior points of the polynomial ghost extrapolation (4 / 6 / 8)
int kernel_max_deriv_table(int kernel);  // highest d for which K^(d) is tabulated
int kernel_max_eval_deriv(int kernel);   // _max_shipped_derivative, get_precomputed_kernel_and_range.jl:33-40
const double *host_table(int kernel, int d);
const double *host_table_i1(int kernel);  // antiderivative table K~

// ---- device-side view of one interpolant (passed by value to kernels) -------------------------
struct ItpParams {
  const void *coefs;       // T*, column-major with pitches below
  long long s1, s2;        // element strides of axis 1 / axis 2
  int ndim;
  int nk[3];               // padded extents = knot vector lengths
  int eqs[3];
  const double *knots[3];  // device
  const double *tabk[3];   // device, K^(d)    101 x 2 eqs
  const double *tabd[3];   // device, K^(d+1)
  int tab_slot_k[3];       // index of the distinct table in the packed smem image
  int tab_slot_d[3];
  int tab_slot_e[3];       // K^(d+2), quintic subgrid only (-1 otherwise)
  int subgrid[3];          // cib200_subgrid: 0 cubic, 1 quintic (1-D / 2-D only), 2 linear
  int prerange_pos;        // 1: eager_2d.jl (linear, linear) position arithmetic (row / t from the stored pre_range)
  int n_tab_slots;
  int npre[3];             // rows of axis d's K table: 101 (shipped), or the generated length of a :linear axis
  int tab_slot_rows[9];    // rows of each distinct table (tables with more than 101 rows stay in global memory)
  int tab_slot_eqs[9];     // eqs of each distinct table
  const double *tab_slot_ptr[9];
  double h[3];
  double lo[3], hi[3];     // _domain_bounds, extrapolate_point.jl:172-176
  double knot0[3];         // knots[d][0] (kernel-parameter constant: never re-loaded inside a loop)
  double scale;            // prod (-1/h_d)^DO_d, associated as the reference does
  double eps_line;         // min(h)/100, extrapolate_point.jl:138-139
  int extrap;
};

// coordinate source: SoA point list or Cartesian product of per-axis vectors
struct Coords {
  const void *x[3];
  long long n[3];  // grid mode: per-axis lengths
  int grid;        // 0: x[d][q], 1: x[d][q_d]
  long long qoff;  // added to q when a failing query is reported (host-chunked calls)
};

struct EvalStatus {
  unsigned long long first_oob;
  unsigned long long first_nan;
};

}  // namespace cib

// opaque handle behind cib200_itp*
struct cib200_itp {
  int device;
  int dtype;
  int eval_mode;
  int perdim;
  int kernel[3], deriv[3];
  long long cshape[3];
  void *d_coefs;
  double *d_knots[3];
  double *d_tables;  // packed distinct tables
  cib::EvalStatus *d_status;
  cib::ItpParams p;
  // staging for host-memory calls (lazily allocated): ring of NSTAGE device buffers, three role streams
  // (copy-in, compute, copy-out) chained by events; all kernels of a call run on ONE stream, so the binned
  // strategy's scratch is never shared between chunks in flight
  static constexpr int NSTAGE = 3;
  void *d_stage[NSTAGE];
  size_t stage_bytes;
  // pinned twin of the ring for PAGEABLE caller buffers (lazily allocated): the caller's arrays are packed into /
  // unpacked from it by a few host threads while the device works on the previous chunk
  void *h_stage[NSTAGE];
  size_t h_stage_bytes;
  cudaStream_t streams[3];          // 0: H2D, 1: compute, 2: D2H
  cudaEvent_t ev_in[NSTAGE], ev_comp[NSTAGE], ev_out[NSTAGE];
  // scratch for the binned strategy (lazily allocated, grows)
  void *d_bin_scratch;
  size_t bin_scratch_bytes;
  void *bin_state;  // opaque, owned by cib200_binned.cu
  // optional pass timing (cib200_itp_set_profiling): events on the launching stream around the passes of a call
  // serialisation of concurrent calls on one handle (cib200_api.cu eval_common)
  std::mutex *mu;
  cudaEvent_t ev_busy;
  int busy_valid;
  int profiling;
  cudaEvent_t prof_ev[5];  // before binning | after scatter | after cell sort | after evaluation | after spill
  int prof_valid;
  int last_strategy;  // cib200_strategy of the most recent evaluation
};

namespace cib {
double jl_pow_int(double x, long long n);  // Julia Base pow_body restated (cib200_eval.cu)
int sm_count();
int prof_mark(cib200_itp *h, int i, cudaStream_t st);  // records prof_ev[i] when profiling is enabled
void *tma_encode_tiled_fn();  // cuTensorMapEncodeTiled through the runtime's driver entry point, or nullptr
// K1 launchers (cib200_eval.cu)
int launch_eval(const cib200_itp *h, const Coords &c, void *out, long long nq, cudaStream_t st);
// binned strategy (cib200_binned.cu)
bool binned_supported(const cib200_itp *h);
long long binned_scratch_bytes(const cib200_itp *h, long long nq);
long long binned_tile_count(const cib200_itp *h);
int launch_eval_binned(cib200_itp *h, const Coords &c, void *out, long long nq, cudaStream_t st);
void binned_destroy(cib200_itp *h);
// sweeps (cib200_sweep.cu)
int sweep_resample(int ndim, int dtype, const long long *n_in, const long long *n_out, const void *d_in,
                   void *d_out, const double *const *knots_in, const double *const *knots_out,
                   const int *kernel, const int *deriv, const int *bcl, const int *bcr, const int *subgrid,
                   const double *const *gen_tables, const int *gen_rows, cudaStream_t st);
int sweep_build_coefs(int ndim, int dtype, const long long *n, const void *d_values, const double *h, const int *kernel,
                      const int *bcl, const int *bcr, void *d_coefs, cudaStream_t st);
// persistent TMA-fed passes of convolution_resample (cib200_resample.cu); -1: the call does not qualify
int resample_pipe_all(int ndim, const long long *n_in, const long long *n_out, const double *d_in, double *d_out,
                      const double *const *knots_in, const double *const *knots_out, const int *kernel, const int *eqs_a,
                      const double *const *tk, const double *const *td, const double *const *te, const int *linear,
                      const int *npre, const double *scale, const int *bcl, const int *bcr, cudaStream_t st);
// persistent TMA-fed FIR pass of convolution_smooth (cib200_smooth.cu); -1: the pass does not qualify
int smooth_pipe_pass(const double *in, double *out, long long inner, long long n, long long outer, int E,
                     const double *kw, cudaStream_t st);
int sweep_smooth(int ndim, int dtype, const long long *n, const void *d_in, void *d_out, const double *h,
                 double B, cudaStream_t st);
}  // namespace cib
