// fdlbm.cu -- host side of the C ABI declared in include/fdlbm.h (engine life cycle, layout
// conversion, step scheduling).  Kernels live in lbm_kernels.cuh / lbm_fused.cuh.
#include "../../include/fdlbm.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "lbm_kernels.cuh"
#include "lbm_fused.cuh"
#include "lbm_fused_vec.cuh"
#include "lbm_init.cuh"

using namespace fdlbm;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return fail(FDLBM_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                                    \
    } while (0)

enum { ST_EMPTY = 0, ST_PRE = 1, ST_POST = 2 };

}  // namespace

struct fdlbm_engine {
    fdlbm_config cfg{};
    int Wl = 0, Hp = 0, ncols = 0;
    size_t esize = 8;
    void *lat[2] = {nullptr, nullptr};
    void *psi[2] = {nullptr, nullptr};
    void *fields = nullptr;  // 9 planes of ncols*Hp reals: rho ux uy p mu mix_tau gx gy lap
    uint8_t *reflect = nullptr, *solid_bytes = nullptr;
    uint32_t *solid = nullptr;
    void *inlet = nullptr, *outlet = nullptr;
    void *staging = nullptr;
    size_t staging_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;                       // host<->device transfers of set_state / get_state
    cudaEvent_t stage_full[2] = {nullptr, nullptr};           // staging buffer b holds a produced plane
    cudaEvent_t stage_free[2] = {nullptr, nullptr};           // ... and has been consumed
    uint64_t stage_seq = 0;
    // fused halo exchange over peer memory (fdlbm_peer_attach)
    uint32_t *flags = nullptr;                                // [0]: low neighbour's progress, [1]: high neighbour's
    void *peer_lat[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [side][lattice]
    uint32_t *peer_flags[2] = {nullptr, nullptr};             // the neighbours' flag words
    int peer_Wl[2] = {0, 0};
    bool peer_ipc[2] = {false, false};
    uint32_t peer_step = 0;                                   // steps taken in peer mode (never reset)
    int cur = 0, pcur = 0;
    int state = ST_EMPTY;
    bool have_geometry = false;
    int64_t iters = 0;
    int64_t launches = 0;
    int kernel = FDLBM_KERNEL_FUSED;

    size_t lat_elems() const { return (size_t)ncols * NPOP * Hp; }
    size_t plane_elems() const { return (size_t)ncols * Hp; }
    bool peer_mode() const { return peer_flags[0] != nullptr || peer_flags[1] != nullptr; }
    bool has_lo() const { return cfg.x0 > 0 || cfg.x_periodic; }
    bool has_hi() const { return cfg.x1 < cfg.W || cfg.x_periodic; }
};

namespace {

template <typename T>
LbmParams<T> make_params(const fdlbm_engine *e, int src, int psrc)
{
    const fdlbm_config &c = e->cfg;
    LbmParams<T> P;
    P.src = (const T *)e->lat[src];
    P.dst = (T *)e->lat[1 - src];
    P.reflect = e->reflect;
    P.solid = e->solid;
    P.psi_old = (const T *)e->psi[psrc];
    P.psi_new = (T *)e->psi[1 - psrc];
    P.inlet_ux = (const T *)e->inlet;
    P.outlet_ux = (const T *)e->outlet;
    P.H = c.H;
    P.Hp = e->Hp;
    P.Wl = e->Wl;
    P.gx0 = c.x0;
    P.W = c.W;
    P.y_wall = c.psi_y_wall;
    P.x_periodic = c.x_periodic;
    P.zou_he = c.zou_he;
    P.inv_tau = (T)(1.0 / c.tau);
    P.gamma = (T)c.gamma;
    P.a = (T)c.a;
    P.kappa = (T)c.kappa;
    P.eta6m = (T)(6.0 * c.Eta_n * c.M);
    P.M = (T)c.M;
    P.psi_wall = (T)c.psi_wall;
    P.psi_left = (T)c.psi_left;
    P.psi_right = (T)c.psi_right;
    P.f3coef = (T)c.outlet_f3_coef;
    P.peer_lo = P.peer_hi = nullptr;
    P.peer_lo_Wl = 0;
    return P;
}

// ---- stream-ordered flags between neighbouring engines (driver entry points fetched at run time so that the
// library keeps loading on machines without libcuda) ---------------------------------------------------------
typedef int (*stream_value32_fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
stream_value32_fn g_write_value = nullptr, g_wait_value = nullptr;

int load_stream_memops()
{
    if (g_write_value && g_wait_value) return 0;
    void *w = nullptr, *q = nullptr;
    cudaDriverEntryPointQueryResult r1, r2;
    CU(cudaGetDriverEntryPoint("cuStreamWriteValue32", &w, cudaEnableDefault, &r1));
    CU(cudaGetDriverEntryPoint("cuStreamWaitValue32", &q, cudaEnableDefault, &r2));
    if (!w || !q || r1 != cudaDriverEntryPointSuccess || r2 != cudaDriverEntryPointSuccess)
        return fail(FDLBM_E_CUDA, "stream memory operations are not available from this driver");
    g_write_value = (stream_value32_fn)w;
    g_wait_value = (stream_value32_fn)q;
    return 0;
}

// wait until both attached neighbours have completed `step` steps (their edge columns of that step are in our
// ghost columns, and they no longer read the lattice our next step writes their ghosts of)
int peer_wait(fdlbm_engine *e, uint32_t step)
{
    for (int side = 0; side < 2; ++side)
        if (e->peer_flags[side]) {
            int rc = g_wait_value(e->stream, (unsigned long long)(uintptr_t)(e->flags + side), step, 0 /* GEQ */);
            if (rc) return fail(FDLBM_E_CUDA, "cuStreamWaitValue32 failed (%d)", rc);
        }
    return 0;
}

// tell the neighbours that this engine has completed `step` steps: we are the HIGH neighbour of our low
// neighbour (its flags[1]) and the LOW neighbour of our high neighbour (its flags[0])
int peer_signal(fdlbm_engine *e, uint32_t step)
{
    for (int side = 0; side < 2; ++side)
        if (e->peer_flags[side]) {
            int rc = g_write_value(e->stream, (unsigned long long)(uintptr_t)(e->peer_flags[side] + (1 - side)), step, 0);
            if (rc) return fail(FDLBM_E_CUDA, "cuStreamWriteValue32 failed (%d)", rc);
        }
    return 0;
}

template <typename T>
void set_peers(const fdlbm_engine *e, LbmParams<T> &P, int dst_lattice)
{
    P.peer_lo = (T *)e->peer_lat[0][dst_lattice];
    P.peer_hi = (T *)e->peer_lat[1][dst_lattice];
    P.peer_lo_Wl = e->peer_Wl[0];
}

template <typename T>
FieldPtrs<T> field_ptrs(const fdlbm_engine *e)
{
    T *b = (T *)e->fields;
    const size_t n = e->plane_elems();
    FieldPtrs<T> F;
    F.rho = b;
    F.ux = b + n;
    F.uy = b + 2 * n;
    F.p = b + 3 * n;
    F.mu = b + 4 * n;
    F.mix_tau = b + 5 * n;
    F.gx = b + 6 * n;
    F.gy = b + 7 * n;
    F.lap = b + 8 * n;
    return F;
}

int ensure_staging(fdlbm_engine *e, size_t bytes)
{
    if (bytes <= e->staging_bytes) return 0;
    if (e->copy_stream) CU(cudaStreamSynchronize(e->copy_stream));
    CU(cudaStreamSynchronize(e->stream));
    if (e->staging) cudaFree(e->staging);
    e->staging = nullptr;
    e->staging_bytes = 0;
    CU(cudaMalloc(&e->staging, bytes));
    e->staging_bytes = bytes;
    return 0;
}

int ensure_fields(fdlbm_engine *e)
{
    if (e->fields) return 0;
    CU(cudaMalloc(&e->fields, 9 * e->plane_elems() * e->esize));
    CU(cudaMemsetAsync(e->fields, 0, 9 * e->plane_elems() * e->esize, e->stream));
    return 0;
}

// 1-D grid of (y tile, column) CTAs, y tiles fastest (decoded by block_cell)
dim3 cell_grid(const fdlbm_engine *e, int ncol) { return dim3((unsigned)((e->cfg.H + TPB - 1) / TPB) * (unsigned)ncol); }

// local periodic wrap of the two ghost columns per side (single slab spanning the whole x range)
int wrap_ghosts(fdlbm_engine *e, void *lat)
{
    const size_t col = (size_t)NPOP * e->Hp * e->esize;
    char *b = (char *)lat;
    CU(cudaMemcpyAsync(b, b + (size_t)e->Wl * col, 2 * col, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemcpyAsync(b + (size_t)(e->Wl + G) * col, b + (size_t)G * col, 2 * col, cudaMemcpyDeviceToDevice,
                       e->stream));
    return 0;
}

template <typename T>
int launch_step(fdlbm_engine *e, bool finalize, const std::function<int()> &between = nullptr)
{
    LbmParams<T> P = make_params<T>(e, e->cur, e->pcur);
    if (e->cfg.x_periodic && !e->cfg.external_halo) {
        int rc = wrap_ghosts(e, e->lat[e->cur]);
        if (rc) return rc;
    }
    const int lo = e->has_lo() ? -1 : 0, hi = e->Wl + (e->has_hi() ? 1 : 0);
    if (!finalize && e->peer_mode()) set_peers(e, P, 1 - e->cur);
    if (finalize || e->kernel == FDLBM_KERNEL_TWOPASS) {
        k_psi<T><<<cell_grid(e, hi - lo), TPB, 0, e->stream>>>(P, lo);
        e->launches += 1;
        if (finalize) {
            // psi is complete after the first pass: its download (transpose on this stream, copy on the copy stream)
            // is queued before the field pass, so the PCIe transfer of psi runs while the fields are computed
            int rc = between ? between() : 0;
            if (rc) return rc;
            k_step_twopass<T, true><<<cell_grid(e, e->Wl), TPB, 0, e->stream>>>(P, field_ptrs<T>(e));
        } else {
            k_step_twopass<T, false><<<cell_grid(e, e->Wl), TPB, 0, e->stream>>>(P, field_ptrs<T>(e));
        }
        e->launches += 1;
    } else {
        int rc = launch_fused_auto<T>(P, e->stream);
        if (rc) return fail(FDLBM_E_CUDA, "fused launch configuration failed (%d)", rc);
        e->launches += 1;
    }
    CU(cudaGetLastError());
    return 0;
}

template <typename T>
int do_steps(fdlbm_engine *e, int n)
{
    if (n <= 0) return 0;
    const bool peer = e->peer_mode();
    int rc;
    if (e->state == ST_PRE) {
        // first collision from the caller's macroscopic arrays, in place on lat[cur]
        LbmParams<T> P = make_params<T>(e, 1 - e->cur, e->pcur);  // dst = lat[cur]
        if (peer) {
            // the neighbours must be done with whatever still reads their lat[cur] ghosts
            if ((rc = peer_wait(e, e->peer_step))) return rc;
            set_peers(e, P, e->cur);
        }
        k_collide_first<T><<<cell_grid(e, e->Wl), TPB, 0, e->stream>>>(P, field_ptrs<T>(e), (const T *)e->psi[e->pcur]);
        CU(cudaGetLastError());
        if (peer && (rc = peer_signal(e, ++e->peer_step))) return rc;
        e->launches += 1;
        e->state = ST_POST;
        e->iters += 1;
        n -= 1;
    }
    for (int k = 0; k < n; ++k) {
        if (peer && (rc = peer_wait(e, e->peer_step))) return rc;
        if ((rc = launch_step<T>(e, false))) return rc;
        if (peer && (rc = peer_signal(e, ++e->peer_step))) return rc;
        e->cur ^= 1;
        e->pcur ^= 1;
        e->iters += 1;
    }
    return 0;
}

// host (H, ncw) window <-> device, for one plane family
struct Overlap {
    int lo, n;  // global columns [lo, lo+n) = window ∩ slab
};
Overlap overlap(const fdlbm_engine *e, int col0, int ncols)
{
    const int lo = col0 > e->cfg.x0 ? col0 : e->cfg.x0;
    const int hi = (col0 + ncols) < e->cfg.x1 ? (col0 + ncols) : e->cfg.x1;
    return Overlap{lo, hi > lo ? hi - lo : 0};
}

template <typename T>
int upload_planes(fdlbm_engine *e, const double *host, int nplanes, int col0, int ncw, T *base, size_t plane_stride,
                  size_t xstride)
{
    // Plane by plane through two staging buffers: the host->device copy of plane k+1 (copy stream) runs
    // while plane k is transposed into the device layout (engine stream); events order the reuse.
    const Overlap ov = overlap(e, col0, ncw);
    if (!ov.n) return 0;
    const int H = e->cfg.H;
    const size_t plane_bytes = (size_t)H * ov.n * sizeof(double);
    int rc = ensure_staging(e, 2 * plane_bytes);
    if (rc) return rc;
    const int xl_lo = ov.lo - e->cfg.x0, xl_hi = xl_lo + ov.n;
    dim3 grid((ov.n + 31) / 32, (H + 31) / 32), block(32, 8);
    for (int k = 0; k < nplanes; ++k) {
        const int b = (int)(e->stage_seq & 1);
        char *st = (char *)e->staging + (size_t)b * plane_bytes;
        const double *src = host + (size_t)k * H * ncw + (ov.lo - col0);
        CU(cudaStreamWaitEvent(e->copy_stream, e->stage_free[b], 0));  // the transpose that last read this buffer
        if (ov.n == ncw)
            CU(cudaMemcpyAsync(st, src, plane_bytes, cudaMemcpyHostToDevice, e->copy_stream));
        else
            CU(cudaMemcpy2DAsync(st, (size_t)ov.n * sizeof(double), src, (size_t)ncw * sizeof(double),
                                 (size_t)ov.n * sizeof(double), (size_t)H, cudaMemcpyHostToDevice, e->copy_stream));
        CU(cudaEventRecord(e->stage_full[b], e->copy_stream));
        CU(cudaStreamWaitEvent(e->stream, e->stage_full[b], 0));
        k_transpose_in<double, T><<<grid, block, 0, e->stream>>>((const double *)st, H, ov.n, ov.lo, base + k * plane_stride,
                                                                  xstride, xl_lo, xl_hi, e->cfg.x0, e->cfg.W, 0);
        CU(cudaEventRecord(e->stage_free[b], e->stream));
        e->launches += 1;
        e->stage_seq += 1;
    }
    CU(cudaGetLastError());
    return 0;
}

template <typename T>
int download_planes(fdlbm_engine *e, double *host, int nplanes, int col0, int ncw, const T *base, size_t plane_stride,
                    size_t xstride)
{
    if (!host) return 0;
    const Overlap ov = overlap(e, col0, ncw);
    if (!ov.n) return 0;
    const int H = e->cfg.H;
    const size_t plane_bytes = (size_t)H * ov.n * sizeof(double);
    int rc = ensure_staging(e, 2 * plane_bytes);
    if (rc) return rc;
    const int xl_lo = ov.lo - e->cfg.x0, xl_hi = xl_lo + ov.n;
    dim3 grid((ov.n + 31) / 32, (H + 31) / 32), block(32, 8);
    for (int k = 0; k < nplanes; ++k) {  // transpose of plane k+1 (engine stream) overlaps the D2H copy of plane k
        const int b = (int)(e->stage_seq & 1);
        char *st = (char *)e->staging + (size_t)b * plane_bytes;
        double *dst = host + (size_t)k * H * ncw + (ov.lo - col0);
        CU(cudaStreamWaitEvent(e->stream, e->stage_free[b], 0));
        k_transpose_out<T, double><<<grid, block, 0, e->stream>>>(base + k * plane_stride, xstride, xl_lo, xl_hi, e->cfg.x0,
                                                                   (double *)st, H, ov.n, ov.lo);
        CU(cudaEventRecord(e->stage_full[b], e->stream));
        CU(cudaStreamWaitEvent(e->copy_stream, e->stage_full[b], 0));
        if (ov.n == ncw)
            CU(cudaMemcpyAsync(dst, st, plane_bytes, cudaMemcpyDeviceToHost, e->copy_stream));
        else
            CU(cudaMemcpy2DAsync(dst, (size_t)ncw * sizeof(double), st, (size_t)ov.n * sizeof(double),
                                 (size_t)ov.n * sizeof(double), (size_t)H, cudaMemcpyDeviceToHost, e->copy_stream));
        CU(cudaEventRecord(e->stage_free[b], e->copy_stream));
        e->launches += 1;
        e->stage_seq += 1;
    }
    CU(cudaGetLastError());
    return 0;
}

// host buffers are borrowed for the duration of an API call: wait for every queued transfer
int drain_transfers(fdlbm_engine *e)
{
    CU(cudaStreamSynchronize(e->copy_stream));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

template <typename T>
int set_state_t(fdlbm_engine *e, int col0, int ncw, const fdlbm_fields *in)
{
    const size_t Hp = e->Hp, n = e->plane_elems();
    int rc = ensure_fields(e);
    if (rc) return rc;
    e->cur = 0;
    e->pcur = 0;
    CU(cudaMemsetAsync(e->lat[0], 0, e->lat_elems() * e->esize, e->stream));
    CU(cudaMemsetAsync(e->lat[1], 0, e->lat_elems() * e->esize, e->stream));
    T *lat = (T *)e->lat[0];
    if ((rc = upload_planes<T>(e, in->f, 9, col0, ncw, lat, Hp, (size_t)NPOP * Hp))) return rc;
    if ((rc = upload_planes<T>(e, in->g, 9, col0, ncw, lat + 9 * Hp, Hp, (size_t)NPOP * Hp))) return rc;
    if ((rc = upload_planes<T>(e, in->psi, 1, col0, ncw, (T *)e->psi[0], 0, Hp))) return rc;
    T *fb = (T *)e->fields;
    const double *srcs[9] = {in->rho, in->ux, in->uy, in->p, in->mu, in->mix_tau, in->nabla_psix, in->nabla_psiy,
                             in->nabla_psi2};
    for (int k = 0; k < 9; ++k) {
        if (!srcs[k]) continue;  // nabla_psi2 is optional on input (recomputed from psi)
        if ((rc = upload_planes<T>(e, srcs[k], 1, col0, ncw, fb + k * n, 0, Hp))) return rc;
    }
    e->state = ST_PRE;
    e->iters = 0;
    return drain_transfers(e);
}

template <typename T>
int get_state_t(fdlbm_engine *e, int col0, int ncw, const fdlbm_fields *out)
{
    const size_t Hp = e->Hp, n = e->plane_elems();
    int rc = ensure_fields(e);
    if (rc) return rc;
    const T *lat, *psi;
    bool psi_done = false;
    if (e->state == ST_PRE) {
        lat = (const T *)e->lat[e->cur];
        psi = (const T *)e->psi[e->pcur];
    } else {
        if (e->peer_mode() && (rc = peer_wait(e, e->peer_step))) return rc;
        const bool only_psi = out->psi && !out->f && !out->g && !out->rho && !out->ux && !out->uy && !out->p && !out->mu &&
                              !out->mix_tau && !out->nabla_psix && !out->nabla_psiy && !out->nabla_psi2;
        if (only_psi) {
            // a psi frame (the drivers' snapshots, fingering.py:565-566): the psi pass alone, on the owned columns
            LbmParams<T> P = make_params<T>(e, e->cur, e->pcur);
            if (e->cfg.x_periodic && !e->cfg.external_halo && (rc = wrap_ghosts(e, e->lat[e->cur]))) return rc;
            k_psi<T><<<cell_grid(e, e->Wl), TPB, 0, e->stream>>>(P, 0);
            CU(cudaGetLastError());
            e->launches += 1;
        } else {  // writes lat[1-cur], psi[1-pcur], fields; no swap
            const T *psi_new = (const T *)e->psi[1 - e->pcur];
            rc = launch_step<T>(e, true, [&]() {
                psi_done = true;
                return download_planes<T>(e, out->psi, 1, col0, ncw, psi_new, 0, Hp);
            });
            if (rc) return rc;
        }
        lat = (const T *)e->lat[1 - e->cur];
        psi = (const T *)e->psi[1 - e->pcur];
    }
    if ((rc = download_planes<T>(e, out->f, 9, col0, ncw, lat, Hp, (size_t)NPOP * Hp))) return rc;
    if ((rc = download_planes<T>(e, out->g, 9, col0, ncw, lat + 9 * Hp, Hp, (size_t)NPOP * Hp))) return rc;
    if (!psi_done && (rc = download_planes<T>(e, out->psi, 1, col0, ncw, psi, 0, Hp))) return rc;
    const T *fb = (const T *)e->fields;
    double *dsts[9] = {out->rho, out->ux, out->uy, out->p, out->mu, out->mix_tau, out->nabla_psix, out->nabla_psiy,
                       out->nabla_psi2};
    for (int k = 0; k < 9; ++k)
        if ((rc = download_planes<T>(e, dsts[k], 1, col0, ncw, fb + k * n, 0, Hp))) return rc;
    return drain_transfers(e);
}

int upload_profile(fdlbm_engine *e, const double *host, void **dev)
{
    const int H = e->cfg.H;
    CU(cudaMalloc(dev, (size_t)e->Hp * e->esize));
    if (e->cfg.dtype == FDLBM_F64) {
        CU(cudaMemcpy(*dev, host, (size_t)H * sizeof(double), cudaMemcpyHostToDevice));
    } else {
        std::vector<float> tmp(host, host + H);
        CU(cudaMemcpy(*dev, tmp.data(), (size_t)H * sizeof(float), cudaMemcpyHostToDevice));
    }
    return 0;
}

template <typename T>
int init_state_t(fdlbm_engine *e, const fdlbm_init *spec)
{
    const fdlbm_config &c = e->cfg;
    int rc = ensure_fields(e);
    if (rc) return rc;
    e->cur = 0;
    e->pcur = 0;
    // k_init_cells writes every population and every field of every owned cell, and the first step every owned cell of
    // the other lattice: only the ghost columns are cleared here (the padding rows are zero since fdlbm_create), not
    // 2 x 2.4 GB of lattice at 8192 x 2048 -- a third of the call
    const size_t ghost = (size_t)G * NPOP * e->Hp * e->esize, body = (size_t)e->Wl * NPOP * e->Hp * e->esize;
    for (int k = 0; k < 2; ++k) {
        CU(cudaMemsetAsync(e->lat[k], 0, ghost, e->stream));
        CU(cudaMemsetAsync((char *)e->lat[k] + ghost + body, 0, ghost, e->stream));
    }
    if (spec->rho) {
        CU(cudaMemsetAsync(e->fields, 0, 9 * e->plane_elems() * e->esize, e->stream));
        if ((rc = upload_planes<T>(e, spec->rho, 1, spec->col0, spec->ncols, (T *)e->fields, 0, (size_t)e->Hp))) return rc;
    }
    InitParams I{};
    I.variant = spec->variant;
    I.n_inject = spec->n_inject;
    I.have_rho = spec->rho != nullptr;
    I.psi_inject = spec->psi_inject, I.psi_rest = spec->psi_rest, I.rho0 = spec->rho0;
    I.gamma = c.gamma, I.a = c.a, I.kappa = c.kappa, I.Eta_n = c.Eta_n, I.M = c.M;
    I.psi_wall = c.psi_wall, I.psi_left = c.psi_left, I.psi_right = c.psi_right;
    LbmParams<T> P = make_params<T>(e, 1, 0);  // dst = lat[0]
    k_init_psi<T><<<cell_grid(e, e->ncols), TPB, 0, e->stream>>>(P, I, (T *)e->psi[0]);
    k_init_cells<T><<<cell_grid(e, e->Wl), TPB, 0, e->stream>>>(P, I, field_ptrs<T>(e));
    CU(cudaGetLastError());
    e->launches += 2;
    e->state = ST_PRE;
    e->iters = 0;
    return drain_transfers(e);
}

}  // namespace

extern "C" {

int fdlbm_abi_version(void) { return FDLBM_ABI_VERSION; }
const char *fdlbm_last_error(void) { return g_err.c_str(); }

int fdlbm_device_count(void)
{
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess) return fail(FDLBM_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(err));
    return n;
}

int fdlbm_create(const fdlbm_config *cfg, fdlbm_engine **out)
{
    if (!cfg || !out) return fail(FDLBM_E_ARG, "null argument");
    *out = nullptr;
    if (cfg->H < 4 || cfg->W < 4) return fail(FDLBM_E_ARG, "grid too small: H=%d W=%d (need >= 4)", cfg->H, cfg->W);
    if (cfg->dtype != FDLBM_F64 && cfg->dtype != FDLBM_F32) return fail(FDLBM_E_ARG, "bad dtype %d", cfg->dtype);
    if (cfg->x0 < 0 || cfg->x1 > cfg->W || cfg->x1 - cfg->x0 < 2)
        return fail(FDLBM_E_ARG, "bad slab [%d,%d) of W=%d (need >= 2 columns)", cfg->x0, cfg->x1, cfg->W);
    if (cfg->zou_he != FDLBM_ZH_NONE && cfg->x_periodic)
        return fail(FDLBM_E_ARG, "Zou-He faces and x-periodic wrap are mutually exclusive");
    if (cfg->zou_he == FDLBM_ZH_NONE && !cfg->x_periodic)
        return fail(FDLBM_E_ARG, "x faces need either Zou-He (FP/FG) or x_periodic (validation)");
    if (cfg->zou_he != FDLBM_ZH_NONE && (!cfg->inlet_ux || !cfg->outlet_ux))
        return fail(FDLBM_E_ARG, "Zou-He faces need inlet_ux and outlet_ux profiles");
    if (cfg->x_periodic && !cfg->external_halo && (cfg->x0 != 0 || cfg->x1 != cfg->W))
        return fail(FDLBM_E_ARG, "a slab of an x-periodic grid needs external_halo=1");
    if (!cfg->x_periodic && !cfg->external_halo && (cfg->x0 != 0 || cfg->x1 != cfg->W))
        return fail(FDLBM_E_ARG, "a proper slab needs external_halo=1");
    if (!(cfg->tau > 0)) return fail(FDLBM_E_ARG, "tau must be positive");
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev == 0)
        return fail(FDLBM_E_CUDA, "no CUDA device (%s); this library has no CPU path",
                    err != cudaSuccess ? cudaGetErrorString(err) : "device count 0");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(FDLBM_E_ARG, "bad device %d of %d", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));

    fdlbm_engine *e = new fdlbm_engine();
    e->cfg = *cfg;
    e->cfg.inlet_ux = e->cfg.outlet_ux = nullptr;  // borrowed pointers are not kept
    e->Wl = cfg->x1 - cfg->x0;
    e->Hp = (cfg->H + 31) / 32 * 32;
    e->ncols = e->Wl + 2 * G;
    e->esize = cfg->dtype == FDLBM_F64 ? 8 : 4;
    e->kernel = cfg->kernel == FDLBM_KERNEL_AUTO ? FDLBM_KERNEL_FUSED : cfg->kernel;
#define CUE(call)                                                                                     \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            fail(_e == cudaErrorMemoryAllocation ? FDLBM_E_NOMEM : FDLBM_E_CUDA, "%s failed: %s", #call, \
                 cudaGetErrorString(_e));                                                             \
            fdlbm_destroy(e);                                                                         \
            return _e == cudaErrorMemoryAllocation ? FDLBM_E_NOMEM : FDLBM_E_CUDA;                    \
        }                                                                                             \
    } while (0)
    CUE(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CUE(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
        CUE(cudaEventCreateWithFlags(&e->stage_full[k], cudaEventDisableTiming));
        CUE(cudaEventCreateWithFlags(&e->stage_free[k], cudaEventDisableTiming));
    }
    for (int k = 0; k < 2; ++k) {
        CUE(cudaMalloc(&e->lat[k], e->lat_elems() * e->esize));
        CUE(cudaMemsetAsync(e->lat[k], 0, e->lat_elems() * e->esize, e->stream));
        CUE(cudaMalloc(&e->psi[k], e->plane_elems() * e->esize));
        CUE(cudaMemsetAsync(e->psi[k], 0, e->plane_elems() * e->esize, e->stream));
    }
    CUE(cudaMalloc((void **)&e->flags, 256));
    CUE(cudaMemsetAsync(e->flags, 0, 256, e->stream));
    // the flag arrays carry one spare (zero) column: the step kernels load flags three columns ahead without a bound check
    CUE(cudaMalloc(&e->reflect, e->plane_elems() + e->Hp));
    CUE(cudaMemsetAsync(e->reflect, 0, e->plane_elems() + e->Hp, e->stream));
    CUE(cudaMalloc(&e->solid_bytes, e->plane_elems()));
    CUE(cudaMemsetAsync(e->solid_bytes, 0, e->plane_elems(), e->stream));
    CUE(cudaMalloc(&e->solid, (e->plane_elems() + e->Hp) / 8));
    CUE(cudaMemsetAsync(e->solid, 0, (e->plane_elems() + e->Hp) / 8, e->stream));
#undef CUE
    if (cfg->zou_he != FDLBM_ZH_NONE) {
        int rc = upload_profile(e, cfg->inlet_ux, &e->inlet);
        if (!rc) rc = upload_profile(e, cfg->outlet_ux, &e->outlet);
        if (rc) {
            fdlbm_destroy(e);
            return rc;
        }
    }
    cudaError_t se = cudaStreamSynchronize(e->stream);
    if (se != cudaSuccess) {
        fdlbm_destroy(e);
        return fail(FDLBM_E_CUDA, "stream sync failed: %s", cudaGetErrorString(se));
    }
    *out = e;
    return 0;
}

void fdlbm_destroy(fdlbm_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (int side = 0; side < 2; ++side)
        if (e->peer_ipc[side]) {
            cudaIpcCloseMemHandle(e->peer_lat[side][0]);
            cudaIpcCloseMemHandle(e->peer_lat[side][1]);
            cudaIpcCloseMemHandle(e->peer_flags[side]);
        }
    void *ptrs[] = {e->lat[0], e->lat[1], e->psi[0], e->psi[1], e->fields, e->reflect, e->solid_bytes,
                    e->solid, e->inlet, e->outlet, e->staging, e->flags};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    for (int k = 0; k < 2; ++k) {
        if (e->stage_full[k]) cudaEventDestroy(e->stage_full[k]);
        if (e->stage_free[k]) cudaEventDestroy(e->stage_free[k]);
    }
    delete e;
}

int fdlbm_set_geometry(fdlbm_engine *e, int col0, int ncols, const uint8_t *solid, const uint8_t *reflect)
{
    if (!e || !solid || !reflect) return fail(FDLBM_E_ARG, "null argument");
    if (ncols <= 0 || col0 < 0 || col0 + ncols > e->cfg.W) return fail(FDLBM_E_ARG, "bad column window");
    // one call carries the whole geometry (the planes are rebuilt from scratch below): every local column that
    // exists in the grid, ghosts included, must be inside the window
    for (int xl = -G; xl < e->Wl + G; ++xl) {
        int gx = e->cfg.x0 + xl;
        if (e->cfg.x_periodic) gx = ((gx % e->cfg.W) + e->cfg.W) % e->cfg.W;
        if (gx < 0 || gx >= e->cfg.W) continue;
        if (gx < col0 || gx >= col0 + ncols)
            return fail(FDLBM_E_ARG, "geometry window [%d,%d) does not cover column %d of the slab [%d,%d) and its ghost columns",
                        col0, col0 + ncols, gx, e->cfg.x0, e->cfg.x1);
    }
    CU(cudaSetDevice(e->cfg.device));
    const int H = e->cfg.H;
    const size_t bytes = (size_t)H * ncols;
    int rc = ensure_staging(e, 2 * bytes);
    if (rc) return rc;
    uint8_t *st = (uint8_t *)e->staging;
    CU(cudaMemcpyAsync(st, solid, bytes, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(st + bytes, reflect, bytes, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemsetAsync(e->solid_bytes, 0, e->plane_elems(), e->stream));
    CU(cudaMemsetAsync(e->reflect, 0, e->plane_elems(), e->stream));
    // all local columns including the ghosts (their flags drive psi on the one-cell ring)
    dim3 grid((e->ncols + 31) / 32, (H + 31) / 32), block(32, 8);
    k_transpose_in<uint8_t, uint8_t><<<grid, block, 0, e->stream>>>(st, H, ncols, col0, e->solid_bytes, (size_t)e->Hp, -G,
                                                                     e->Wl + G, e->cfg.x0, e->cfg.W, e->cfg.x_periodic);
    k_transpose_in<uint8_t, uint8_t><<<grid, block, 0, e->stream>>>(st + bytes, H, ncols, col0, e->reflect, (size_t)e->Hp,
                                                                     -G, e->Wl + G, e->cfg.x0, e->cfg.W, e->cfg.x_periodic);
    const size_t n = e->plane_elems();
    k_pack_solid<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->solid_bytes, e->solid, n);
    e->launches += 3;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    e->have_geometry = true;
    return 0;
}

int fdlbm_set_state(fdlbm_engine *e, int col0, int ncols, const fdlbm_fields *in)
{
    if (!e || !in) return fail(FDLBM_E_ARG, "null argument");
    if (!in->f || !in->g || !in->psi || !in->rho || !in->ux || !in->uy || !in->p || !in->mu || !in->mix_tau ||
        !in->nabla_psix || !in->nabla_psiy)
        return fail(FDLBM_E_ARG, "set_state needs f, g, psi, rho, ux, uy, p, mu, mix_tau, nabla_psix, nabla_psiy");
    if (ncols <= 0 || col0 < 0 || col0 + ncols > e->cfg.W) return fail(FDLBM_E_ARG, "bad column window");
    if (col0 > e->cfg.x0 || col0 + ncols < e->cfg.x1)  // both lattices are cleared below: a partial window would lose the rest
        return fail(FDLBM_E_ARG, "state window [%d,%d) does not cover the slab [%d,%d)", col0, col0 + ncols, e->cfg.x0, e->cfg.x1);
    if (!e->have_geometry) return fail(FDLBM_E_STATE, "set_geometry must be called before set_state");
    CU(cudaSetDevice(e->cfg.device));
    return e->cfg.dtype == FDLBM_F64 ? set_state_t<double>(e, col0, ncols, in) : set_state_t<float>(e, col0, ncols, in);
}

int fdlbm_init_state(fdlbm_engine *e, const fdlbm_init *spec)
{
    if (!e || !spec) return fail(FDLBM_E_ARG, "null argument");
    if (spec->variant != FDLBM_INIT_FP && spec->variant != FDLBM_INIT_FG) return fail(FDLBM_E_ARG, "bad init variant %d", spec->variant);
    if (spec->n_inject < 0) return fail(FDLBM_E_ARG, "negative n_inject");
    if (!spec->rho && !(spec->rho0 > 0)) return fail(FDLBM_E_ARG, "rho0 must be positive");
    if (spec->rho && (spec->ncols <= 0 || spec->col0 > e->cfg.x0 || spec->col0 + spec->ncols < e->cfg.x1))
        return fail(FDLBM_E_ARG, "the rho window [%d,%d) does not cover the slab [%d,%d)", spec->col0, spec->col0 + spec->ncols,
                    e->cfg.x0, e->cfg.x1);
    if (!e->have_geometry) return fail(FDLBM_E_STATE, "set_geometry must be called before init_state");
    CU(cudaSetDevice(e->cfg.device));
    return e->cfg.dtype == FDLBM_F64 ? init_state_t<double>(e, spec) : init_state_t<float>(e, spec);
}

int fdlbm_step(fdlbm_engine *e, int n)
{
    if (!e) return fail(FDLBM_E_ARG, "null engine");
    if (n < 0) return fail(FDLBM_E_ARG, "negative step count");
    if (e->state == ST_EMPTY) return fail(FDLBM_E_STATE, "set_state must be called before step");
    if (e->cfg.external_halo && n > 1 && !e->peer_mode())
        return fail(FDLBM_E_ARG, "external_halo engines advance one step per call (halo exchange in between) "
                                 "unless their neighbours are attached with fdlbm_peer_attach");
    CU(cudaSetDevice(e->cfg.device));
    return e->cfg.dtype == FDLBM_F64 ? do_steps<double>(e, n) : do_steps<float>(e, n);
}

int fdlbm_get_state(fdlbm_engine *e, int col0, int ncols, const fdlbm_fields *out)
{
    if (!e || !out) return fail(FDLBM_E_ARG, "null argument");
    if (e->state == ST_EMPTY) return fail(FDLBM_E_STATE, "no state loaded");
    if (ncols <= 0 || col0 < 0 || col0 + ncols > e->cfg.W) return fail(FDLBM_E_ARG, "bad column window");
    CU(cudaSetDevice(e->cfg.device));
    return e->cfg.dtype == FDLBM_F64 ? get_state_t<double>(e, col0, ncols, out) : get_state_t<float>(e, col0, ncols, out);
}

int64_t fdlbm_iterations(const fdlbm_engine *e) { return e ? e->iters : -1; }

int fdlbm_sync(fdlbm_engine *e)
{
    if (!e) return fail(FDLBM_E_ARG, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

void *fdlbm_stream(fdlbm_engine *e) { return e ? (void *)e->stream : nullptr; }
int64_t fdlbm_launch_count(const fdlbm_engine *e) { return e ? e->launches : -1; }

int fdlbm_halo_regions(fdlbm_engine *e, fdlbm_halo *out)
{
    if (!e || !out) return fail(FDLBM_E_ARG, "null argument");
    const size_t col = (size_t)NPOP * e->Hp * e->esize;
    char *b = (char *)e->lat[e->cur];
    out->recv_lo = b;
    out->send_lo = b + (size_t)G * col;
    out->send_hi = b + (size_t)e->Wl * col;
    out->recv_hi = b + (size_t)(e->Wl + G) * col;
    out->bytes = (size_t)G * col;
    return 0;
}

// checkpoint blob: header | lat[cur] | psi[pcur] | fields (9 planes)
struct CkptHeader {
    char magic[8];
    int32_t H, W, x0, x1, dtype, state, Hp, ncols;
    int64_t iters;
};

size_t fdlbm_checkpoint_bytes(const fdlbm_engine *e)
{
    if (!e) return 0;
    return sizeof(CkptHeader) + (e->lat_elems() + 10 * e->plane_elems()) * e->esize;
}

int fdlbm_checkpoint_save(fdlbm_engine *e, void *host, size_t bytes)
{
    if (!e || !host) return fail(FDLBM_E_ARG, "null argument");
    if (e->state == ST_EMPTY) return fail(FDLBM_E_STATE, "no state loaded");
    if (bytes < fdlbm_checkpoint_bytes(e)) return fail(FDLBM_E_ARG, "checkpoint buffer too small");
    CU(cudaSetDevice(e->cfg.device));
    int rc = ensure_fields(e);
    if (rc) return rc;
    CkptHeader h{};
    memcpy(h.magic, "FDLBMCK1", 8);
    h.H = e->cfg.H, h.W = e->cfg.W, h.x0 = e->cfg.x0, h.x1 = e->cfg.x1, h.dtype = e->cfg.dtype;
    h.state = e->state, h.Hp = e->Hp, h.ncols = e->ncols, h.iters = e->iters;
    char *p = (char *)host;
    memcpy(p, &h, sizeof h);
    p += sizeof h;
    // the ghost columns of lat[cur] are written by the attached neighbours' step kernels: wait for their step
    // `peer_step` like get_state does (all ranks of a slab run checkpoint / restore at the same step)
    if (e->peer_mode() && (rc = peer_wait(e, e->peer_step))) return rc;
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(p, e->lat[e->cur], e->lat_elems() * e->esize, cudaMemcpyDeviceToHost));
    p += e->lat_elems() * e->esize;
    CU(cudaMemcpy(p, e->psi[e->pcur], e->plane_elems() * e->esize, cudaMemcpyDeviceToHost));
    p += e->plane_elems() * e->esize;
    CU(cudaMemcpy(p, e->fields, 9 * e->plane_elems() * e->esize, cudaMemcpyDeviceToHost));
    return 0;
}

int fdlbm_checkpoint_load(fdlbm_engine *e, const void *host, size_t bytes)
{
    if (!e || !host) return fail(FDLBM_E_ARG, "null argument");
    if (bytes < fdlbm_checkpoint_bytes(e)) return fail(FDLBM_E_ARG, "checkpoint blob too small for this engine");
    if (!e->have_geometry) return fail(FDLBM_E_STATE, "set_geometry must be called before loading a checkpoint");
    CkptHeader h;
    memcpy(&h, host, sizeof h);
    if (memcmp(h.magic, "FDLBMCK1", 8) != 0) return fail(FDLBM_E_ARG, "not a checkpoint blob");
    if (h.H != e->cfg.H || h.W != e->cfg.W || h.x0 != e->cfg.x0 || h.x1 != e->cfg.x1 || h.dtype != e->cfg.dtype ||
        h.Hp != e->Hp || h.ncols != e->ncols)
        return fail(FDLBM_E_ARG, "checkpoint was taken from a different grid / slab / dtype");
    if (h.state != ST_PRE && h.state != ST_POST) return fail(FDLBM_E_ARG, "checkpoint blob holds no state (%d)", h.state);
    CU(cudaSetDevice(e->cfg.device));
    int rc = ensure_fields(e);
    if (rc) return rc;
    CU(cudaStreamSynchronize(e->stream));
    const char *p = (const char *)host + sizeof h;
    e->cur = 0;
    e->pcur = 0;
    CU(cudaMemcpy(e->lat[0], p, e->lat_elems() * e->esize, cudaMemcpyHostToDevice));
    p += e->lat_elems() * e->esize;
    CU(cudaMemset(e->lat[1], 0, e->lat_elems() * e->esize));
    CU(cudaMemcpy(e->psi[0], p, e->plane_elems() * e->esize, cudaMemcpyHostToDevice));
    p += e->plane_elems() * e->esize;
    CU(cudaMemcpy(e->fields, p, 9 * e->plane_elems() * e->esize, cudaMemcpyHostToDevice));
    e->state = h.state;
    e->iters = h.iters;
    return 0;
}

int fdlbm_count_nonfinite(fdlbm_engine *e, int64_t *n_bad)
{
    if (!e || !n_bad) return fail(FDLBM_E_ARG, "null argument");
    if (e->state == ST_EMPTY) return fail(FDLBM_E_STATE, "no state loaded");
    CU(cudaSetDevice(e->cfg.device));
    unsigned long long *d = (unsigned long long *)(e->flags + 16);  // spare words of the flag block
    CU(cudaMemsetAsync(d, 0, sizeof *d, e->stream));
    if (e->cfg.dtype == FDLBM_F64)
        k_count_nonfinite<double><<<cell_grid(e, e->Wl), TPB, 0, e->stream>>>((const double *)e->lat[e->cur], e->cfg.H, e->Hp, d);
    else
        k_count_nonfinite<float><<<cell_grid(e, e->Wl), TPB, 0, e->stream>>>((const float *)e->lat[e->cur], e->cfg.H, e->Hp, d);
    CU(cudaGetLastError());
    e->launches += 1;
    unsigned long long h = 0;
    CU(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    *n_bad = (int64_t)h;
    return 0;
}

int fdlbm_peer_export(fdlbm_engine *e, fdlbm_peer_info *out)
{
    if (!e || !out) return fail(FDLBM_E_ARG, "null argument");
    CU(cudaSetDevice(e->cfg.device));
    memset(out, 0, sizeof *out);
    out->pid = (int64_t)getpid();
    out->device = e->cfg.device;
    out->Wl = e->Wl;
    out->Hp = e->Hp;
    out->dtype = e->cfg.dtype;
    out->H = e->cfg.H;
    out->lat[0] = e->lat[0];
    out->lat[1] = e->lat[1];
    out->flags = e->flags;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)out->ipc_lat[0], e->lat[0]));
    CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)out->ipc_lat[1], e->lat[1]));
    CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)out->ipc_flags, e->flags));
    return 0;
}

int fdlbm_peer_attach(fdlbm_engine *e, int side, const fdlbm_peer_info *nb)
{
    if (!e || !nb || (side != 0 && side != 1)) return fail(FDLBM_E_ARG, "bad argument");
    if (!e->cfg.external_halo) return fail(FDLBM_E_ARG, "peer halos need an engine created with external_halo=1");
    if (nb->H != e->cfg.H || nb->Hp != e->Hp || nb->dtype != e->cfg.dtype)
        return fail(FDLBM_E_ARG, "neighbour has a different H or dtype");
    if (e->peer_flags[side]) return fail(FDLBM_E_STATE, "side %d is already attached", side);
    CU(cudaSetDevice(e->cfg.device));
    int rc = load_stream_memops();
    if (rc) return rc;
    if (nb->pid == (int64_t)getpid()) {  // same process: the pointers are directly usable
        if (nb->device != e->cfg.device) {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, e->cfg.device, nb->device));
            if (!can) return fail(FDLBM_E_CUDA, "device %d cannot access device %d", e->cfg.device, nb->device);
            cudaError_t pe = cudaDeviceEnablePeerAccess(nb->device, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
                return fail(FDLBM_E_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(pe));
            cudaGetLastError();
        }
        e->peer_lat[side][0] = nb->lat[0];
        e->peer_lat[side][1] = nb->lat[1];
        e->peer_flags[side] = (uint32_t *)nb->flags;
    } else {
        void *p0 = nullptr, *p1 = nullptr, *pf = nullptr;
        CU(cudaIpcOpenMemHandle(&p0, *(const cudaIpcMemHandle_t *)nb->ipc_lat[0], cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle(&p1, *(const cudaIpcMemHandle_t *)nb->ipc_lat[1], cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle(&pf, *(const cudaIpcMemHandle_t *)nb->ipc_flags, cudaIpcMemLazyEnablePeerAccess));
        e->peer_lat[side][0] = p0;
        e->peer_lat[side][1] = p1;
        e->peer_flags[side] = (uint32_t *)pf;
        e->peer_ipc[side] = true;
    }
    e->peer_Wl[side] = nb->Wl;
    return 0;
}

void *fdlbm_pinned_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        fail(FDLBM_E_NOMEM, "cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}

void fdlbm_pinned_free(void *p)
{
    if (p) cudaFreeHost(p);
}

}  // extern "C"

#include "lbm_ops.cuh"
