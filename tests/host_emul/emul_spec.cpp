// emul_spec.cpp — HOST build of a run-time specialised objective (TEST INFRASTRUCTURE ONLY).
//
// alphabeta-rs_b200/csrc/abfit_jit.cu turns a pedigree's micro-op program into straight-line CUDA C++ and compiles
// it with NVRTC for the GPU.  The generated function has no cross-lane communication, so the very same source
// (abfit_jit_dump writes it to a file) can be compiled for the CPU with the shims of this directory and compared
// with the oracle bit for bit — the `-m "not gpu"` suite checks the code generator without a GPU.
//
//   g++ ... -DSPEC_SOURCE='"/path/to/generated.cu"' -include shims.h emul_spec.cpp     (tests/test_host_emul.py)
#include <cmath>
#include <cstring>
#include <vector>

#include "shims.h"
#define ABFIT_HOST_EMUL 1
#include SPEC_SOURCE

using namespace abfit;

namespace {
WarpCtx make_ctx(const abfit_problem *pb, std::vector<double> &D)
{
    D.resize((size_t)pb->n_pairs + 4);
    for (int i = 0; i < pb->n_pairs; ++i) D[i] = pb->pedigree[4 * (size_t)i + 3];
    WarpCtx c;
    std::memset(&c, 0, sizeof c);
    c.D = D.data();
    c.n_pairs = pb->n_pairs;
    c.p_uu0 = pb->p0uu;
    c.p_mm0 = 1.0 - pb->p0uu;
    c.eqp = pb->eqp;
    c.penw = pb->eqp_weight * (double)pb->n_pairs;
    return c;
}
}  // namespace

extern "C" {

int spec_cost(const abfit_problem *pb, const double *theta, int B, double *cost, double *lse)
{
    std::vector<double> D;
    const WarpCtx c = make_ctx(pb, D);
    const DBroadcast Dat{c.D};
    for (int i = 0; i < B; ++i) {
        const double *t = theta + 4 * (size_t)i;
        cost[i] = SpecObjective::eval(c, Dat, i & 31, t[0], t[1], t[2], t[3], true);
        if (lse) lse[i] = SpecObjective::eval(c, Dat, i & 31, t[0], t[1], t[2], t[3], false);
    }
    return 0;
}

// one Nelder-Mead run per simplex, the state machine driven as the kernels drive it; dstar: optional per-fit D*
// columns [n][n_pairs] read through the per-lane column accessor (bootstrap replicates)
int spec_fit(const abfit_problem *pb, const double *simplices, int n, const double *dstar, int max_iters, double sd_tol,
             uint32_t flags, double var_thr, double range_thr, abfit_fit *out)
{
    std::vector<double> D;
    const WarpCtx c = make_ctx(pb, D);
    NMParams nm{max_iters, sd_tol, flags, var_thr, range_thr};
    std::vector<double> simplex(25 * 32), tile;
    for (int f = 0; f < n; ++f) {
        const int lane = f & 31;
        LaneSimplex S;
        S.X = simplex.data() + lane;
        S.C = S.X + 20 * 32;
        if (dstar) {
            tile.assign((size_t)c.n_pairs * 32, 0.0);
            for (int i = 0; i < c.n_pairs; ++i) tile[(size_t)i * 32 + lane] = dstar[(size_t)f * c.n_pairs + i];
        }
        LaneNM L;
        std::memset(&L, 0, sizeof L);
        for (int q = 0; q < 20; ++q) S.X[q * 32] = simplices[(size_t)f * 20 + q];
        nm_begin(L, S, f);
        abfit_fit res;
        std::memset(&res, 0, sizeof res);
        for (;;) {
            double v;
            if (dstar) {
                const DLaneColumn Dat{tile.data() + lane};
                v = SpecObjective::eval(c, Dat, lane, L.xt[0], L.xt[1], L.xt[2], L.xt[3], L.phase != PH_LSE);
            } else {
                const DBroadcast Dat{c.D};
                v = SpecObjective::eval(c, Dat, lane, L.xt[0], L.xt[1], L.xt[2], L.xt[3], L.phase != PH_LSE);
            }
            if (nm_advance(L, S, nm, v, res, 1u << lane)) break;
        }
        out[f] = res;
    }
    return 0;
}
}
