/*
 * fdlbm.h -- C ABI of the B200-native two-phase D2Q9 lattice-Boltzmann engine.
 *
 * The reference (ebinan92/Fingering_dynamics) has no FFI: its boundary is the Python module surface
 * of lattice_boltzmann/*.py, and its hot path is the body of `for i in range(MAX_T)` inlined in each
 * driver's main() (fingering_periodic.py:454-479, fingering.py:558-585, validation.py:392-409).
 * This header is what a ctypes binding of that path binds instead.  Every entry
 * point names the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; host arrays are borrowed for the duration of the call;
 *   - host arrays use the REFERENCE layout: C-order float64, populations (9,H,ncols), fields (H,ncols),
 *     x (axis 1) = flow direction; `col0,ncols` say which global columns the arrays hold
 *     (col0=0, ncols=W for the whole grid);
 *   - every function returns 0 on success or a negative FDLBM_E_* code and never throws; the message is
 *     available from fdlbm_last_error(); there is NO CPU fallback: without a CUDA device
 *     fdlbm_create fails with FDLBM_E_CUDA;
 *   - an engine owns its device memory and one CUDA stream; calls on one engine must be serialised by
 *     the caller, distinct engines are independent.
 */
#ifndef FDLBM_H
#define FDLBM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDLBM_ABI_VERSION 1

enum { FDLBM_OK = 0, FDLBM_E_ARG = -1, FDLBM_E_CUDA = -2, FDLBM_E_STATE = -3, FDLBM_E_NOMEM = -4 };
enum { FDLBM_F64 = 0, FDLBM_F32 = 1 };
/* Zou-He flavour of the x faces.
 * FDLBM_ZH_FP: fingering_periodic.py:268-324 (all rows, per-row velocity profile on both faces);
 * FDLBM_ZH_FG: fingering.py:298-390 (inlet rows 1..H-2 + the two corner nodes, outlet corner copies). */
enum { FDLBM_ZH_NONE = 0, FDLBM_ZH_FP = 1, FDLBM_ZH_FG = 2 };
/* step kernel: FUSED = one pass per step (psi ring recomputed on chip);
 * TWOPASS = psi staged through HBM, kept as the in-library cross-check of the fused kernel. */
enum { FDLBM_KERNEL_AUTO = 0, FDLBM_KERNEL_TWOPASS = 1, FDLBM_KERNEL_FUSED = 2 };

typedef struct fdlbm_engine fdlbm_engine;

typedef struct {
    int32_t H, W;        /* global grid: rows (y) x columns (x)                                        */
    int32_t dtype;       /* FDLBM_F64 | FDLBM_F32: storage and arithmetic type on the device           */
    int32_t psi_y_wall;  /* 0: psi stencil wraps in y (fingering_periodic.py:218-225);
                            1: ghost rows = psi_wall (fingering.py:224, validation.py:246)             */
    int32_t x_periodic;  /* 0: ghost columns psi_left/psi_right (fingering_periodic.py:94-95,218);
                            1: everything wraps in x (validation.py)                                   */
    int32_t zou_he;      /* FDLBM_ZH_*                                                                 */
    int32_t kernel;      /* FDLBM_KERNEL_*                                                             */
    int32_t device;      /* CUDA device ordinal                                                        */
    int32_t x0, x1;      /* slab: global columns [x0,x1) owned by this engine (0,W = whole grid)       */
    int32_t external_halo; /* 1: the caller fills the 2 ghost columns per side before every step
                              (multi-GPU slabs, see fdlbm_halo_regions); 0: the engine wraps locally   */
    double tau;          /* relaxation time of g (fingering_periodic.py:26,263)                        */
    double gamma;        /* mobility (fingering_periodic.py:40,164,168)                                */
    double a, kappa;     /* mu = a psi (1-psi^2) - kappa lap psi (fingering_periodic.py:141-144);
                            validation.py's a psi (psi^2-1) with a>0 is the same with -a               */
    double Eta_n, M;     /* viscosities of tau_mix (fingering_periodic.py:201-208)                     */
    double psi_wall;     /* wettability: psi of solid cells and ghost rows (fingering_periodic.py:93,212) */
    double psi_left, psi_right; /* psi ghost columns and Zou-He targets (+1 / -1)                      */
    double outlet_f3_coef;      /* 2/3 (fingering_periodic.py:317) or 1.5 (fingering.py:378)           */
    const double *inlet_ux;     /* H per-row face velocities (fingering_periodic.py:270-271), or NULL  */
    const double *outlet_ux;    /* H values (fingering_periodic.py:306-307), or NULL                   */
} fdlbm_config;

/* Host-side view of the macroscopic state the reference's Compute object holds
 * (fingering_periodic.py:98-113).  "Masked" 1-D reference arrays are passed as full (H,ncols) grids;
 * entries at solid cells are ignored on input and zero on output.  Any pointer may be NULL on output. */
typedef struct {
    double *f, *g;                          /* (9,H,ncols) */
    double *psi, *rho, *ux, *uy, *p, *mu, *mix_tau;
    double *nabla_psix, *nabla_psiy, *nabla_psi2;
} fdlbm_fields;

int fdlbm_abi_version(void);
/* message of the last failing call on this thread (also valid when fdlbm_create failed) */
const char *fdlbm_last_error(void);
/* number of CUDA devices visible, or FDLBM_E_CUDA */
int fdlbm_device_count(void);

/* Replaces Createblock + Bounce_back + Compute construction as the owner of the run's state. */
int fdlbm_create(const fdlbm_config *cfg, fdlbm_engine **out);
void fdlbm_destroy(fdlbm_engine *e);

/* Geometry.  solid: 1 = block cell (block_psi_all == 1, fingering_periodic.py:450-451).
 * reflect: bit (i-1) set <=> after streaming, f_i and g_i of this cell are replaced by the pre-stream
 * f_opp(i), g_opp(i) of the SAME cell -- the union of the class tables of bounce_back.py:89-167 /
 * 25-86 and of the wall rows (fingering.py:573, validation.py:357-376).  Both (H,ncols) uint8.
 * ONE call carries the whole geometry of the engine: the window [col0, col0+ncols) must cover the slab AND its two
 * ghost columns per side where those lie inside the grid (wrapped around for x-periodic grids), i.e.
 * [max(0,x0-2), min(W,x1+2)) at least; earlier geometry is discarded.  FDLBM_E_ARG otherwise. */
int fdlbm_set_geometry(fdlbm_engine *e, int col0, int ncols, const uint8_t *solid, const uint8_t *reflect);

/* Load the state a reference iteration starts from: populations f,g plus the macroscopic arrays the
 * FIRST collision reads (rho, ux, uy, p, mu, mix_tau, psi, nabla_psix, nabla_psiy) exactly as the
 * caller holds them -- Compute.__init__ leaves them mutually inconsistent (fingering_periodic.py:111,
 * fingering.py:121-123) and parity needs that.  Resets the step counter and discards the previous state: the
 * window [col0, col0+ncols) must cover the slab [x0,x1) (FDLBM_E_ARG otherwise); ghost columns are not taken
 * from it (they come from the halo exchange / local wrap). */
int fdlbm_set_state(fdlbm_engine *e, int col0, int ncols, const fdlbm_fields *in);

/* Compute.__init__ ON THE DEVICE (fingering_periodic.py:90-121, fingering.py:95-127) -- the alternative to
 * fdlbm_set_state for grids whose 29 host planes would be the job's dominant transfer: psi = psi_inject on the
 * first n_inject GLOBAL columns and psi_rest elsewhere (psi_wall on solids), rho = rho0 (or the caller's plane:
 * fingering.py:106-107 draws it at random), u, mu, p, tau_mix, grad psi and f = f_eq, g = g_eq in the reference's
 * operation order with the quirks parity depends on (FP: mu is still 0 everywhere, fingering_periodic.py:111;
 * FG: p before mu, uy from mu but not ux, fingering.py:121-123).  fp64 engines get the reference's bits.
 * Leaves the engine where fdlbm_set_state would (first collision from these arrays); resets the step counter. */
enum { FDLBM_INIT_FP = 1, FDLBM_INIT_FG = 2 };
typedef struct {
    int32_t variant;              /* FDLBM_INIT_*                                                    */
    int32_t n_inject;             /* 5 in both drivers (fingering_periodic.py:91, fingering.py:96)   */
    double psi_inject, psi_rest;  /* +1 / -1                                                         */
    double rho0;                  /* used when rho == NULL (fingering_periodic.py:104: 1.0)          */
    const double *rho;            /* optional (H,ncols) host plane holding global columns [col0, col0+ncols)
                                     which must cover the slab; entries at solid cells are ignored   */
    int32_t col0, ncols;
} fdlbm_init;
int fdlbm_init_state(fdlbm_engine *e, const fdlbm_init *spec);

/* Advance n reference iterations (fingering_periodic.py:455-479).  Asynchronous on the engine stream. */
int fdlbm_step(fdlbm_engine *e, int n);

/* Read back what the reference holds after the iterations done so far: pre-collision f, g and all
 * macroscopic fields (fingering_periodic.py:470-479).  Does not disturb the run. */
int fdlbm_get_state(fdlbm_engine *e, int col0, int ncols, const fdlbm_fields *out);

/* iterations completed since fdlbm_set_state */
int64_t fdlbm_iterations(const fdlbm_engine *e);
/* block until everything queued on the engine stream is done */
int fdlbm_sync(fdlbm_engine *e);
/* the engine's cudaStream_t (for CUDA-event timing on the launching stream) */
void *fdlbm_stream(fdlbm_engine *e);
/* kernels launched by this engine since creation (bench.py's gpu_launches) */
int64_t fdlbm_launch_count(const fdlbm_engine *e);

/* Multi-GPU slabs (no counterpart in the reference, which is single-process).  Device pointers into
 * the lattice holding the current state: per side, `bytes` contiguous bytes to send (the 2 owned
 * edge columns) and to receive into (the 2 ghost columns).  Valid until the next fdlbm_step. */
typedef struct {
    void *send_lo, *recv_lo, *send_hi, *recv_hi;
    size_t bytes;
} fdlbm_halo;
int fdlbm_halo_regions(fdlbm_engine *e, fdlbm_halo *out);

/* Exact restart (absent in the reference, whose runs keep everything in RAM): the raw device state --
 * current lattice with ghosts, psi, macroscopic scratch, counters -- as an opaque blob.  A run continued
 * from a loaded checkpoint is bit-identical to the uninterrupted run.  Geometry is not part of the blob.
 * Peer-attached slabs: every engine of the run saves / loads at the SAME step, and a barrier between the loads and
 * the next fdlbm_step keeps a neighbour's halo stores out of a lattice that is still being restored. */
size_t fdlbm_checkpoint_bytes(const fdlbm_engine *e);
int fdlbm_checkpoint_save(fdlbm_engine *e, void *host, size_t bytes);
int fdlbm_checkpoint_load(fdlbm_engine *e, const void *host, size_t bytes);

/* Watchdog standing in for np.seterr(all='raise') (fingering_periodic.py:497): counts non-finite
 * populations of the owned columns of the current state into *n_bad. */
int fdlbm_count_nonfinite(fdlbm_engine *e, int64_t *n_bad);

/* Halo exchange FUSED into the step (preferred on NVLink): once the neighbours are attached, the step
 * kernel stores its two edge columns straight into the neighbours' ghost columns through peer-mapped
 * memory, and the steps of neighbouring engines are ordered by stream-side flags (cuStreamWriteValue32 /
 * cuStreamWaitValue32 on peer memory) -- no host synchronisation, no collective library on the data path,
 * and fdlbm_step(e, n) may advance many steps per call.  Every engine of a run must take the same
 * sequence of set_state / step calls.
 *   fdlbm_peer_export: describe this engine (CUDA IPC handles of its two lattices and its flag words,
 *                      plus raw pointers for engines living in the same process);
 *   fdlbm_peer_attach: side 0 = `nb` owns the columns just below x0, side 1 = just above x1. */
typedef struct {
    int64_t pid;                /* process that owns the memory */
    int32_t device, Wl, Hp, dtype, H;
    void *lat[2], *flags;       /* valid inside process `pid` */
    unsigned char ipc_lat[2][64], ipc_flags[64]; /* cudaIpcMemHandle_t */
} fdlbm_peer_info;
int fdlbm_peer_export(fdlbm_engine *e, fdlbm_peer_info *out);
int fdlbm_peer_attach(fdlbm_engine *e, int side, const fdlbm_peer_info *nb);

/* page-locked host memory for the e2e path */
void *fdlbm_pinned_alloc(size_t bytes);
void fdlbm_pinned_free(void *p);

/* ---- stateless operators: NumPy-in / NumPy-out twins of the reference's module functions -------- */
/* Stateless for the caller: every call takes host arrays and returns host arrays.  Inside, the device buffers of an
 * operator call are kept for the next call with the same configuration (at most four configurations; a call holds a
 * process-wide lock, the operators are not re-entrant); fdlbm_op_release frees them. */
void fdlbm_op_release(void);
/* stream(f, g), fingering_periodic.py:327-343 (in place, all cells, wrap on both axes) */
int fdlbm_op_stream(int H, int W, double *f, double *g);
/* Bounce_back.halfway_bounceback_* / bottom_top_wall with the classes already folded into reflect bits */
int fdlbm_op_bounce_back(int H, int W, const uint8_t *reflect, const double *f_behind, const double *g_behind,
                         double *f, double *g);
/* Compute.getNabla_psix/psiy/psi2 (fingering_periodic.py:214-256 and twins); outputs may be NULL */
int fdlbm_op_stencils(const fdlbm_config *cfg, const double *psi, double *gx, double *gy, double *lap);
/* the collision of one iteration on fluid cells (fingering_periodic.py:455-460), in place on f, g */
int fdlbm_op_collide(const fdlbm_config *cfg, const uint8_t *solid, const fdlbm_fields *io);
/* Compute.getfeq / getgeq / getLarge_F for all nine directions (fingering_periodic.py:171-199) from the
 * macroscopic arrays in `in`; feq, geq, F are (9,H,W) outputs (zero on solid cells), any may be NULL */
int fdlbm_op_collision_terms(const fdlbm_config *cfg, const uint8_t *solid, const fdlbm_fields *in, double *feq,
                             double *geq, double *F);
/* The point-wise getters of Compute from the arrays in `in` (psi, rho, mu, nabla_psi2), each output (H,W), any
 * may be NULL: p = getP (fingering_periodic.py:123-124, uses in->mu), mu = getMu_plain (:146-149, uses
 * in->nabla_psi2), mix_tau = getMix_tau (:201-208), a0/a1_8/b0/b1_8 = getA0/getA1_8/getB0/getB1_8 (:155-169,
 * use in->p and in->mu). */
typedef struct {
    double *p, *mu, *mix_tau, *a0, *a1_8, *b0, *b1_8;
} fdlbm_algebra_out;
int fdlbm_op_algebra(const fdlbm_config *cfg, const fdlbm_fields *in, const fdlbm_algebra_out *out);
/* zou_he_boundary_inlet + _outlet (fingering_periodic.py:268-324 / fingering.py:298-390), in place */
int fdlbm_op_zou_he(const fdlbm_config *cfg, const fdlbm_fields *io);
/* the moment updates of one iteration (fingering_periodic.py:470-479): f,g in; all fields out */
int fdlbm_op_moments(const fdlbm_config *cfg, const uint8_t *solid, const fdlbm_fields *io);

#ifdef __cplusplus
}
#endif
#endif /* FDLBM_H */
