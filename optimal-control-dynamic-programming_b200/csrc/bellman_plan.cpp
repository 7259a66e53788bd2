// bellman_plan.cpp — host-only logic: descriptor validation, locate-mode choice, exact reach
// analysis and slab planning.  Nothing here touches CUDA, so it is usable (and tested) on a
// machine without a GPU.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "bellman_internal.h"

namespace bellman {

static bool grid_is_uniform(const double *s, int n) {
    const double h = (s[n - 1] - s[0]) / (double)(n - 1);
    for (int i = 0; i < n; ++i) {
        const double dev = std::fabs(s[i] - (s[0] + (double)i * h));
        if (!(dev <= 1e-14 * (s[n - 1] - s[0]))) return false;
    }
    return true;
}

std::string load_problem(const bellman_desc *d, HostProblem &hp) {
    if (!d) return "descriptor is NULL";
    if (d->struct_size != (int32_t)sizeof(bellman_desc))
        return "bellman_desc.struct_size mismatch (ABI version " +
               std::to_string(BELLMAN_ABI_VERSION) + ")";
    if (d->D < 2 || d->D > MAXD) return "D must be 2.." + std::to_string(MAXD);
    if (d->C < 1) return "C must be >= 1";
    if (d->P < 1) return "P must be >= 1";
    if (d->N < 2) return "N must be >= 2";
    if (!d->r) return "r table is NULL";
    hp.D = d->D; hp.C = d->C; hp.P = d->P; hp.N = d->N;
    hp.idx_bytes = d->idx_bytes == 0 ? 4 : d->idx_bytes;
    if (hp.idx_bytes != 1 && hp.idx_bytes != 2 && hp.idx_bytes != 4) return "idx_bytes must be 0, 1, 2 or 4";
    if ((hp.idx_bytes == 1 && d->C > 256) || (hp.idx_bytes == 2 && d->C > 65536)) return "idx_bytes too small for C controls";
    hp.part_cuts.clear();
    if (d->part_cuts && d->nranks >= 1) hp.part_cuts.assign(d->part_cuts, d->part_cuts + d->nranks + 1);
    for (int k = 0; k < hp.D; ++k) {
        if (d->n[k] < 2) return "every grid needs >= 2 points";
        hp.n[k] = d->n[k];
    }
    bool seen[MAXD] = {false, false, false, false};
    for (int k = 0; k < hp.D; ++k) {
        const int o = d->q_order[k];
        if (o < 0 || o >= hp.D || seen[o]) return "q_order must be a permutation of 0..D-1";
        seen[o] = true;
        hp.q_order[k] = o;
    }
    hp.mode.assign((size_t)hp.P * hp.D, BELLMAN_LOCATE_UNIFORM);
    for (int k = 0; k < hp.D; ++k) {
        const int n = hp.n[k];
        if (!d->grid[k] || !d->Ta[k] || !d->q[k]) return "grid/Ta/q table is NULL";
        if (d->src_a[k] < 0 || d->src_a[k] >= hp.D) return "src_a out of range";
        hp.src_a[k] = d->src_a[k];
        hp.has_b[k] = d->Tb[k] != nullptr && d->src_b[k] >= 0;
        if (hp.has_b[k] && d->src_b[k] >= hp.D) return "src_b out of range";
        hp.src_b[k] = hp.has_b[k] ? d->src_b[k] : -1;
        hp.has_c[k] = d->Tc[k] != nullptr;
        hp.grid[k].assign(d->grid[k], d->grid[k] + (size_t)hp.P * n);
        hp.q[k].assign(d->q[k], d->q[k] + (size_t)hp.P * n);
        hp.Ta[k].assign(d->Ta[k], d->Ta[k] + (size_t)hp.P * hp.n[hp.src_a[k]]);
        if (hp.has_b[k]) hp.Tb[k].assign(d->Tb[k], d->Tb[k] + (size_t)hp.P * hp.n[hp.src_b[k]]);
        if (hp.has_c[k]) hp.Tc[k].assign(d->Tc[k], d->Tc[k] + (size_t)hp.P * hp.C);
        hp.rinv[k].assign((size_t)hp.P * n, 0.0);
        hp.inv_h[k].assign(hp.P, 0.0);
        hp.off[k].assign(hp.P, 0.0);
        for (int p = 0; p < hp.P; ++p) {
            const double *s = hp.grid[k].data() + (size_t)p * n;
            for (int i = 0; i + 1 < n; ++i) {
                if (!(s[i + 1] > s[i])) return "grid vectors must be strictly increasing";
                hp.rinv[k][(size_t)p * n + i] = 1.0 / (s[i + 1] - s[i]);
            }
            hp.inv_h[k][p] = (double)(n - 1) / (s[n - 1] - s[0]);
            hp.off[k][p] = -(s[0] * hp.inv_h[k][p]);
            hp.mode[(size_t)p * hp.D + k] =
                grid_is_uniform(s, n) ? BELLMAN_LOCATE_UNIFORM : BELLMAN_LOCATE_SEARCH;
        }
    }
    // SEARCH dimensions: bucket table (does not change the result of the bin search, only where it starts)
    for (int k = 0; k < hp.D; ++k) {
        const int n = hp.n[k];
        hp.lut_n[k] = 4 * n;
        hp.lut[k].assign((size_t)hp.P * (hp.lut_n[k] + 1), 0);
        hp.lut_invw[k].assign(hp.P, 0.0);
        for (int p = 0; p < hp.P; ++p) {
            const double *s = hp.grid[k].data() + (size_t)p * n;
            const double w = (s[n - 1] - s[0]) / (double)hp.lut_n[k];
            hp.lut_invw[k][p] = 1.0 / w;
            int cell = 0;
            for (int b = 0; b <= hp.lut_n[k]; ++b) {
                const double edge = s[0] + (double)b * w;
                while (cell < n - 2 && s[cell + 1] <= edge) ++cell;
                hp.lut[k][(size_t)p * (hp.lut_n[k] + 1) + b] = cell;
            }
        }
    }
    // UNIFORM dimensions are evaluated in cell units (include/bellman.h): rescale the next-state
    // tables of every (problem, dimension) that uses the uniform rule, one rounding per entry
    for (int k = 0; k < hp.D; ++k)
        for (int p = 0; p < hp.P; ++p) {
            if (hp.mode[(size_t)p * hp.D + k] != BELLMAN_LOCATE_UNIFORM) continue;
            const double ih = hp.inv_h[k][p], of = hp.off[k][p];
            const int na = hp.n[hp.src_a[k]];
            for (int i = 0; i < na; ++i) {
                double &v = hp.Ta[k][(size_t)p * na + i];
                v = std::fma(v, ih, of);
            }
            if (hp.has_b[k]) {
                const int nb = hp.n[hp.src_b[k]];
                for (int i = 0; i < nb; ++i) hp.Tb[k][(size_t)p * nb + i] *= ih;
            }
            if (hp.has_c[k])
                for (int c = 0; c < hp.C; ++c) hp.Tc[k][(size_t)p * hp.C + c] *= ih;
        }
    hp.r.assign(d->r, d->r + (size_t)hp.P * hp.C);
    for (double v : hp.r)
        if (!std::isfinite(v)) return "r table must be finite";
    return "";
}

// x is a query in the units the kernel sees: the fractional cell coordinate for UNIFORM
// dimensions, the state value for SEARCH dimensions
int host_locate(const HostProblem &hp, int p, int d, double x) {
    const int n = hp.n[d];
    const double *s = hp.grid[d].data() + (size_t)p * n;
    int cell;
    if (hp.mode[(size_t)p * hp.D + d] == BELLMAN_LOCATE_UNIFORM) {
        if (!(x >= 0.0)) cell = 0;
        else if (x >= (double)(n - 1)) cell = n - 2;
        else cell = (int)x;
    } else {
        cell = (int)(std::upper_bound(s, s + n, x) - s) - 1;
        cell = std::min(std::max(cell, 0), n - 2);
    }
    return cell;
}

static void minmax(const double *v, int lo, int hi, double &mn, double &mx) {
    mn = std::numeric_limits<double>::infinity();
    mx = -mn;
    for (int i = lo; i < hi; ++i) { mn = std::min(mn, v[i]); mx = std::max(mx, v[i]); }
}

// Floating-point addition is monotone in each argument, so the extreme queries are the sums of
// the extreme table entries, formed with the kernel's own association: (Ta + Tb) + Tc.
void reach_range(const HostProblem &hp, int dim, int own_lo, int own_hi, int &ext_lo, int &ext_hi) {
    ext_lo = own_lo;
    ext_hi = own_hi;
    for (int p = 0; p < hp.P; ++p) {
        const int sa = hp.src_a[dim];
        double amn, amx;
        minmax(hp.Ta[dim].data() + (size_t)p * hp.n[sa], sa == dim ? own_lo : 0,
               sa == dim ? own_hi : hp.n[sa], amn, amx);
        double lo = amn, hi = amx;
        if (hp.has_b[dim]) {
            const int sb = hp.src_b[dim];
            double bmn, bmx;
            minmax(hp.Tb[dim].data() + (size_t)p * hp.n[sb], sb == dim ? own_lo : 0,
                   sb == dim ? own_hi : hp.n[sb], bmn, bmx);
            lo = lo + bmn;
            hi = hi + bmx;
        }
        if (hp.has_c[dim]) {
            double cmn, cmx;
            minmax(hp.Tc[dim].data() + (size_t)p * hp.C, 0, hp.C, cmn, cmx);
            lo = lo + cmn;
            hi = hi + cmx;
        }
        ext_lo = std::min(ext_lo, host_locate(hp, p, dim, lo));
        ext_hi = std::max(ext_hi, host_locate(hp, p, dim, hi) + 2);
    }
}

std::string plan_slabs(const HostProblem &hp, int part_dim, int nranks, bellman_slab *out) {
    if (part_dim < 0 || part_dim >= hp.D) return "part_dim out of range";
    if (nranks < 1) return "nranks must be >= 1";
    const int n = hp.n[part_dim];
    if (nranks > n) return "more ranks than grid points along part_dim";
    const bool cuts = !hp.part_cuts.empty();
    if (cuts) {
        if ((int)hp.part_cuts.size() != nranks + 1 || hp.part_cuts[0] != 0 || hp.part_cuts[nranks] != n)
            return "part_cuts must hold nranks + 1 boundaries from 0 to n[part_dim]";
        for (int r = 0; r < nranks; ++r)
            if (hp.part_cuts[r + 1] <= hp.part_cuts[r]) return "part_cuts must be strictly increasing";
    }
    for (int r = 0; r < nranks; ++r) {
        const int lo = cuts ? hp.part_cuts[r] : (int)((int64_t)n * r / nranks);
        const int hi = cuts ? hp.part_cuts[r + 1] : (int)((int64_t)n * (r + 1) / nranks);
        out[r].own_lo = lo;
        out[r].own_hi = hi;
        int elo, ehi;
        // the slab restricts only the index along part_dim; other dimensions range fully, and
        // only dimension part_dim of the query decides which slabs are read
        reach_range(hp, part_dim, lo, hi, elo, ehi);
        out[r].ext_lo = std::max(0, elo);
        out[r].ext_hi = std::min(n, ehi);
    }
    return "";
}

}  // namespace bellman

// ---------------------------------------------------------------------------------------------
// C ABI (host-only entry points)
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_plan_error;
const char *bellman_plan_last_error() { return g_plan_error.c_str(); }

extern "C" int bellman_query_locate(const bellman_desc *d, int32_t *mode_out) {
    bellman::HostProblem hp;
    g_plan_error = bellman::load_problem(d, hp);
    if (!g_plan_error.empty() || !mode_out) return BELLMAN_ERR_BAD_ARG;
    std::memcpy(mode_out, hp.mode.data(), sizeof(int32_t) * hp.mode.size());
    return BELLMAN_OK;
}

extern "C" int bellman_plan_slabs(const bellman_desc *d, int32_t part_dim, int32_t nranks,
                                  bellman_slab *slabs_out) {
    bellman::HostProblem hp;
    g_plan_error = bellman::load_problem(d, hp);
    if (!g_plan_error.empty() || !slabs_out) return BELLMAN_ERR_BAD_ARG;
    g_plan_error = bellman::plan_slabs(hp, part_dim, nranks, slabs_out);
    return g_plan_error.empty() ? BELLMAN_OK : BELLMAN_ERR_BAD_ARG;
}
