/* cfx_oracle.cpp -- CPU ORACLE for the charge-flux Ewald path. TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A double-precision, single-threaded restatement of the plugin's Reference-platform kernel
 * (/root/reference/platforms/reference/src/ReferenceCoulKernels.cpp) and of the three OpenMM helpers
 * it calls (getDeltaR / getDeltaRPeriodic / computeNeighborListVoxelHash, restated from their
 * documented semantics -- OpenMM is a third-party dependency that is not vendored, version unpinned,
 * API era 7.3-7.5).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library. The product (libcfx_b200.so) never links or calls it.
 *
 * Parity pin: the reference ships no tests or golden vectors, so this restatement is pinned against
 * the reference's OWN sources executed here: oracle/_ref/libcfx_ref.so is built from the unmodified
 * files under /root/reference (against the OpenMM stand-in in shim/), and tests/test_oracle_vs_ref.py
 * requires agreement to ~1e-13 relative on energies, forces, charges and neighbour sets. The golden
 * fixtures in tests/golden/ were generated from that reference build (tests/golden/make_golden.py).
 *
 * Build: g++ -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile). -ffp-contract=off keeps the
 * arithmetic free of fused multiply-adds so the in-cutoff predicate is reproducible.
 */
#include "../include/cfx_b200.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <set>
#include <string>
#include <utility>
#include <vector>

namespace {

struct V3 {
    double x, y, z;
    double  operator[](int i) const { return (&x)[i]; }
    double& operator[](int i)       { return (&x)[i]; }
};

struct Oracle {
    int n = 0;
    std::vector<double> q0, halfSigma, twoSqrtEps;          // :237-239
    std::vector<int> bondIdx, angleIdx, waterIdx;
    std::vector<double> bondPar, anglePar, waterPar;
    int nb = 0, na = 0, nw = 0;
    std::vector<int> rowDq, rowDx;                           // Jacobian COO index tables, :286-383
    std::vector<double> rowVal;                              // [3P]
    std::vector<std::set<int> > excl;                        // :385-391
    bool pbc = false;
    double cutoff = 0, tol = 0, alpha = 0;
    int kmax[3] = {0, 0, 0};
    int kxLo = 0, kxHi = -1;                                 // sampling window for the CPU baseline
    // outputs of the last evaluation
    std::vector<double> q, dedq;
    std::vector<std::pair<int,int> > pairs;
    long long pairCandidates = 0;
};

thread_local std::string g_err;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

/* OpenMM ReferenceForce::getDeltaR: J - I. */
inline V3 deltaPlain(const V3& I, const V3& J) { return V3{J.x-I.x, J.y-I.y, J.z-I.z}; }

/* OpenMM ReferenceForce::getDeltaRPeriodic: subtract c*floor(dz/cz+0.5), then b, then a. */
inline V3 deltaPeriodic(const V3& I, const V3& J, const double* box) {
    V3 d = deltaPlain(I, J);
    for (int axis = 2; axis >= 0; axis--) {
        const double* v = box + 3*axis;
        double s = floor(d[axis]/v[axis] + 0.5);
        d.x -= v[0]*s; d.y -= v[1]*s; d.z -= v[2]*s;
    }
    return d;
}

inline double norm2(const V3& d) { return d.x*d.x + d.y*d.y + d.z*d.z; }

/* ReferenceCoulKernels.cpp:32-35 -- OpenMM's Ewald reciprocal error estimate. */
double ewaldErrorEstimate(int kmax, double width, double alpha) {
    double t = kmax*M_PI/(width*alpha);
    return 0.05*sqrt(width*alpha)*kmax*exp(-t*t);
}

/* ReferenceCoulKernels.cpp:37-228 -- charges q(x) and the sparse Jacobian rows dq/dx.
 * Row order per term is fixed by the index tables built in create(): bond 4 rows, angle/water 9. */
void assembleCharges(Oracle& o, const V3* pos, const double* box) {
    o.q.assign(o.q0.begin(), o.q0.end());
    auto delta = [&](int a, int b) { return o.pbc ? deltaPeriodic(pos[a], pos[b], box) : deltaPlain(pos[a], pos[b]); };

    for (int t = 0; t < o.nb; t++) {                                    // :42-80
        int p1 = o.bondIdx[2*t], p2 = o.bondIdx[2*t+1];
        double k = o.bondPar[2*t], b = o.bondPar[2*t+1];
        V3 d = delta(p1, p2);
        double r = sqrt(norm2(d));
        double dq = k*(r - b);
        o.q[p1] += dq;
        o.q[p2] -= dq;
        double c = k/r;
        double* rows = &o.rowVal[3*4*t];
        for (int j = 0; j < 3; j++) {
            double v = c*d[j];
            rows[0+j] = -v;  rows[3+j] = v;  rows[6+j] = v;  rows[9+j] = -v;
        }
    }
    for (int t = 0; t < o.na; t++) {                                    // :81-162
        int p1 = o.angleIdx[3*t], p2 = o.angleIdx[3*t+1], p3 = o.angleIdx[3*t+2];
        double k = o.anglePar[2*t], theta0 = o.anglePar[2*t+1];
        V3 d21 = delta(p2, p1), d23 = delta(p2, p3), d13 = delta(p1, p3);
        double r21s = norm2(d21), r23s = norm2(d23), r13s = norm2(d13);
        double r21 = sqrt(r21s), r23 = sqrt(r23s);
        double cost = (r23s + r21s - r13s)/2/r21/r23;
        double dq = k*(acos(cost) - theta0);
        o.q[p1] += dq;
        o.q[p3] += dq;
        o.q[p2] -= 2*dq;
        double invRR = 1.0/r21/r23;
        double invSin = 1/sqrt(1 - cost*cost);
        double c1 = k*invRR*invSin;
        double c21 = k*cost*invSin/r21s;
        double c23 = k*cost*invSin/r23s;
        double* rows = &o.rowVal[3*(4*o.nb + 9*t)];
        for (int j = 0; j < 3; j++) {
            double v1 = -c1*d23[j] + c21*d21[j];
            double v3 = -c1*d21[j] + c23*d23[j];
            double v2 = -v1 - v3;
            rows[ 0+j] = v1;     rows[ 3+j] = v2;     rows[ 6+j] = v3;
            rows[ 9+j] = -2*v1;  rows[12+j] = -2*v2;  rows[15+j] = -2*v3;
            rows[18+j] = v1;     rows[21+j] = v2;     rows[24+j] = v3;
        }
    }
    for (int t = 0; t < o.nw; t++) {                                    // :163-227
        int p1 = o.waterIdx[3*t], p2 = o.waterIdx[3*t+1], p3 = o.waterIdx[3*t+2];
        const double* w = &o.waterPar[5*t];
        double k1 = w[0], k2 = w[1], kub = w[2], b0 = w[3], ub0 = w[4];
        V3 d12 = delta(p1, p2), d13 = delta(p1, p3), d23 = delta(p2, p3);
        double r12 = sqrt(norm2(d12)), r13 = sqrt(norm2(d13)), r23 = sqrt(norm2(d23));
        double dq2 = k1*(r12 - b0) + k2*(r13 - b0) + kub*(r23 - ub0);
        double dq3 = k1*(r13 - b0) + k2*(r12 - b0) + kub*(r23 - ub0);
        double dq1 = -dq2 - dq3;
        o.q[p1] += dq1;
        o.q[p2] += dq2;
        o.q[p3] += dq3;
        // OpenMM's Vec3::operator/ multiplies by the reciprocal.
        double i12 = 1.0/r12, i13 = 1.0/r13, i23 = 1.0/r23;
        double* rows = &o.rowVal[3*(4*o.nb + 9*o.na + 9*t)];
        for (int j = 0; j < 3; j++) {
            double n12 = d12[j]*i12, n13 = d13[j]*i13, n23 = d23[j]*i23;
            double a1 = k1*n12, a2 = k2*n12, b1 = k1*n13, b2 = k2*n13, u = kub*n23;
            rows[ 0+j] = a1 + a2 + b1 + b2;
            rows[ 3+j] = -a1 - a2 + 2*u;
            rows[ 6+j] = -b2 - b1 - 2*u;
            rows[ 9+j] = -a1 - b2;
            rows[12+j] = a1 - u;
            rows[15+j] = b2 + u;
            rows[18+j] = -a2 - b1;
            rows[21+j] = a2 - u;
            rows[24+j] = b1 + u;
        }
    }
}

/* OpenMM computeNeighborListVoxelHash, restated: all i<j, not excluded, periodic r2 <= rc2.
 * Enumeration order here matches shim/openmm/shim_runtime.cpp so that the oracle and the reference
 * build sum pair terms in the same order. */
void buildNeighborList(Oracle& o, const V3* pos, const double* box) {
    o.pairs.clear();
    o.pairCandidates = 0;
    const int n = o.n;
    if (n < 2) return;
    const double rc2 = o.cutoff*o.cutoff;
    int nc[3]; double edge[3];
    for (int d = 0; d < 3; d++) {
        double len = box[4*d];
        nc[d] = std::min(256, std::max(1, (int) floor(len/o.cutoff)));
        edge[d] = len/nc[d];
    }
    const int ncells = nc[0]*nc[1]*nc[2];
    std::vector<int> cellOf(n), start(ncells+1, 0), order(n);
    for (int i = 0; i < n; i++) {
        int c[3];
        for (int d = 0; d < 3; d++) {
            double len = box[4*d];
            double x = pos[i][d] - 0.0;
            x -= floor(x/len)*len;
            c[d] = std::max(0, std::min(nc[d]-1, (int) floor(x/edge[d])));
        }
        cellOf[i] = (c[0]*nc[1] + c[1])*nc[2] + c[2];
        start[cellOf[i]+1]++;
    }
    for (int c = 0; c < ncells; c++) start[c+1] += start[c];
    std::vector<int> fill(start.begin(), start.end()-1);
    for (int i = 0; i < n; i++) order[fill[cellOf[i]]++] = i;
    std::vector<int> nbr;
    for (int cx = 0; cx < nc[0]; cx++)
    for (int cy = 0; cy < nc[1]; cy++)
    for (int cz = 0; cz < nc[2]; cz++) {
        nbr.clear();
        for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++)
        for (int dz = -1; dz <= 1; dz++) {
            int a = (cx+dx+nc[0])%nc[0], b = (cy+dy+nc[1])%nc[1], c = (cz+dz+nc[2])%nc[2];
            nbr.push_back((a*nc[1] + b)*nc[2] + c);
        }
        std::sort(nbr.begin(), nbr.end());
        nbr.erase(std::unique(nbr.begin(), nbr.end()), nbr.end());
        int c0 = (cx*nc[1] + cy)*nc[2] + cz;
        for (int a = start[c0]; a < start[c0+1]; a++) {
            int i = order[a];
            const std::set<int>& ex = o.excl[i];
            for (int other : nbr)
                for (int b = start[other]; b < start[other+1]; b++) {
                    int j = order[b];
                    if (j <= i) continue;
                    o.pairCandidates++;
                    double r2 = norm2(deltaPeriodic(pos[i], pos[j], box));
                    if (r2 > rc2) continue;
                    if (ex.count(j)) continue;
                    o.pairs.push_back(std::make_pair(i, j));
                }
        }
    }
}

/* ReferenceCoulKernels.cpp:436-499 -- non-periodic branch: all pairs, minus excluded pairs, chain rule. */
double evaluateNoCutoff(Oracle& o, const V3* pos, bool incF, bool incE, double* energy, V3* F) {
    const double ke = CFX_ONE_4PI_EPS0;
    double e = 0.0, eAll = 0.0, eEx = 0.0;
    auto pairTerm = [&](int i, int j, double sign, double& eAcc) {
        V3 d = deltaPlain(pos[i], pos[j]);
        double invR = 1.0/sqrt(norm2(d));
        double sig = o.halfSigma[i] + o.halfSigma[j];
        double s2 = invR*sig; s2 *= s2;
        double s6 = s2*s2*s2;
        double eps = o.twoSqrtEps[i]*o.twoSqrtEps[j];
        double es6 = s6*eps;
        if (incE) {
            double a = ke*o.q[i]*o.q[j]*invR, b = es6*(s6 - 1);
            if (sign > 0) { e += a; e += b; } else { e -= a; e -= b; }
            eAcc += sign*(a + b);
        }
        if (incF) {
            double dEdR = (es6*(12*s6 - 6) + ke*o.q[i]*o.q[j]*invR)*invR*invR;
            for (int c = 0; c < 3; c++) {
                if (sign > 0) { F[i][c] -= dEdR*d[c]; F[j][c] += dEdR*d[c]; }
                else          { F[i][c] += dEdR*d[c]; F[j][c] -= dEdR*d[c]; }
            }
            if (sign > 0) { o.dedq[i] += ke*o.q[j]*invR; o.dedq[j] += ke*o.q[i]*invR; }
            else          { o.dedq[i] -= ke*o.q[j]*invR; o.dedq[j] -= ke*o.q[i]*invR; }
        }
    };
    for (int i = 0; i < o.n; i++)
        for (int j = i+1; j < o.n; j++)
            pairTerm(i, j, +1.0, eAll);
    for (int i = 0; i < o.n; i++)
        for (int j : o.excl[i])
            if (i < j)
                pairTerm(i, j, -1.0, eEx);
    energy[CFX_E_SELF] = 0; energy[CFX_E_RECIP] = 0;
    energy[CFX_E_DIRECT] = eAll; energy[CFX_E_EXCL] = eEx; energy[CFX_E_TOTAL] = e;
    return e;
}

/* ReferenceCoulKernels.cpp:501-633 -- periodic branch. */
double evaluateEwald(Oracle& o, const V3* pos, const double* box, bool incF, bool incE, double* energy, V3* F) {
    const double ke = CFX_ONE_4PI_EPS0;
    const double alpha = o.alpha, invAlpha2 = 1.0/alpha/alpha;
    const int n = o.n;
    double eSelf = 0, eRecip = 0, eDirect = 0, eExcl = 0;
    for (int i = 0; i < n; i++) {                                        // :507-510
        eSelf -= ke*o.q[i]*o.q[i]*alpha/sqrt(M_PI);
        o.dedq[i] += -2*ke*alpha/sqrt(M_PI)*o.q[i];
    }
    // reciprocal space, :513-556. Only the diagonal of the box enters.
    const double gx = 2*M_PI/box[0], gy = 2*M_PI/box[4], gz = 2*M_PI/box[8];
    const double C = 4.0/box[0]/box[4]/box[8]*M_PI*ke;
    const int hiX = (o.kxHi < 0) ? o.kmax[0] : std::min(o.kxHi, o.kmax[0]);
    for (int nx = std::max(0, o.kxLo); nx < hiX; nx++) {
        double kx = nx*gx;
        for (int ny = (nx == 0 ? 0 : 1 - o.kmax[1]); ny < o.kmax[1]; ny++) {
            double ky = ny*gy;
            for (int nz = ((nx == 0 && ny == 0) ? 1 : 1 - o.kmax[2]); nz < o.kmax[2]; nz++) {
                double kz = nz*gz;
                double k2 = kx*kx + ky*ky + kz*kz;
                double ak = exp(-k2*0.25*invAlpha2)/k2;
                double ss = 0.0, cs = 0.0;
                if (incF || incE)
                    for (int i = 0; i < n; i++) {
                        double gr = kx*pos[i].x + ky*pos[i].y + kz*pos[i].z;
                        cs += o.q[i]*cos(gr);
                        ss += o.q[i]*sin(gr);
                    }
                if (incF)
                    for (int i = 0; i < n; i++) {
                        double gr = kx*pos[i].x + ky*pos[i].y + kz*pos[i].z;
                        double g = 2.0*C*ak*(ss*o.q[i]*cos(gr) - cs*o.q[i]*sin(gr));
                        F[i].x -= g*kx;
                        F[i].y -= g*ky;
                        F[i].z -= g*kz;
                        o.dedq[i] += 2*C*ak*(cs*cos(gr) + ss*sin(gr));
                    }
                if (incE)
                    eRecip += C*ak*(cs*cs + ss*ss);
            }
        }
    }
    // direct space over the neighbour list, :559-593. delta = pos[i] - pos[j] (min image).
    buildNeighborList(o, pos, box);
    for (const auto& pr : o.pairs) {
        int i = pr.first, j = pr.second;
        V3 d = deltaPeriodic(pos[j], pos[i], box);
        double r = sqrt(norm2(d));
        double invR = 1.0/r;
        double ar = alpha*r;
        double sig = o.halfSigma[i] + o.halfSigma[j];
        double s2 = invR*sig; s2 *= s2;
        double s6 = s2*s2*s2;
        double eps = o.twoSqrtEps[i]*o.twoSqrtEps[j];
        double es6 = s6*eps;
        if (incF) {
            double dEdR = ke*o.q[i]*o.q[j]*invR*invR*invR;
            dEdR = dEdR*(erfc(ar) + ar*exp(-ar*ar)*2.0/sqrt(M_PI));
            dEdR += es6*(12*s6 - 6)*invR*invR;
            for (int c = 0; c < 3; c++) {
                double f = dEdR*d[c];
                F[i][c] += f;
                F[j][c] -= f;
            }
            o.dedq[i] += ke*o.q[j]*invR*erfc(ar);
            o.dedq[j] += ke*o.q[i]*invR*erfc(ar);
        }
        eDirect += ke*o.q[i]*o.q[j]*invR*erfc(ar) + es6*(s6 - 1);
    }
    // excluded pairs: remove the reciprocal-space image of the pair (erf term), no cutoff, :596-622.
    for (int i = 0; i < n; i++)
        for (int j : o.excl[i]) {
            if (!(i < j)) continue;
            V3 d = deltaPeriodic(pos[j], pos[i], box);
            double r = sqrt(norm2(d));
            double invR = 1.0/r;
            double ar = alpha*r;
            if (incF) {
                double dEdR = ke*o.q[i]*o.q[j]*invR*invR*invR;
                dEdR = dEdR*(erf(ar) - ar*exp(-ar*ar)*2.0/sqrt(M_PI));
                for (int c = 0; c < 3; c++) {
                    double f = dEdR*d[c];
                    F[i][c] -= f;
                    F[j][c] += f;
                }
                o.dedq[i] -= ke*o.q[j]*invR*erf(ar);
                o.dedq[j] -= ke*o.q[i]*invR*erf(ar);
            }
            eExcl -= ke*o.q[i]*o.q[j]*invR*erf(ar);
        }
    energy[CFX_E_SELF] = eSelf; energy[CFX_E_RECIP] = eRecip;
    energy[CFX_E_DIRECT] = eDirect; energy[CFX_E_EXCL] = eExcl;
    energy[CFX_E_TOTAL] = eSelf + eRecip + eDirect + eExcl;              // :633
    return energy[CFX_E_TOTAL];
}

} // namespace

struct cfxo_handle { Oracle o; };

extern "C" {

const char* cfxo_last_error(void) { return g_err.c_str(); }

/* ReferenceCoulKernels.cpp:230-422 */
int cfxo_create(const cfx_system_desc* d, cfxo_handle** out) {
    if (!d || !out) return fail(CFX_ERR_ARGUMENT, "null argument");
    const int n = d->num_particles;
    if (n < 0) return fail(CFX_ERR_ARGUMENT, "negative particle count");
    cfxo_handle* h = new cfxo_handle();
    Oracle& o = h->o;
    o.n = n;
    o.q0.assign(d->charge, d->charge + n);
    o.halfSigma.resize(n); o.twoSqrtEps.resize(n);
    for (int i = 0; i < n; i++) {
        o.halfSigma[i] = 0.5*d->sigma[i];
        o.twoSqrtEps[i] = 2.0*sqrt(d->epsilon[i]);
    }
    o.nb = d->num_flux_bonds; o.na = d->num_flux_angles; o.nw = d->num_flux_waters;
    o.bondIdx.assign(d->flux_bond_idx, d->flux_bond_idx + 2*o.nb);
    o.bondPar.assign(d->flux_bond_params, d->flux_bond_params + 2*o.nb);
    o.angleIdx.assign(d->flux_angle_idx, d->flux_angle_idx + 3*o.na);
    o.anglePar.assign(d->flux_angle_params, d->flux_angle_params + 2*o.na);
    o.waterIdx.assign(d->flux_water_idx, d->flux_water_idx + 3*o.nw);
    o.waterPar.assign(d->flux_water_params, d->flux_water_params + 5*o.nw);
    auto checkIdx = [&](const std::vector<int>& v) { for (int a : v) if (a < 0 || a >= n) return false; return true; };
    if (!checkIdx(o.bondIdx) || !checkIdx(o.angleIdx) || !checkIdx(o.waterIdx)) {
        delete h; return fail(CFX_ERR_ARGUMENT, "flux term particle index out of range");
    }
    // Jacobian index tables: bond rows (p1,p1),(p1,p2),(p2,p1),(p2,p2); 3-body rows row-major (dq,dx).
    for (int t = 0; t < o.nb; t++)
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) {
            o.rowDq.push_back(o.bondIdx[2*t+a]); o.rowDx.push_back(o.bondIdx[2*t+b]);
        }
    for (int t = 0; t < o.na; t++)
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
            o.rowDq.push_back(o.angleIdx[3*t+a]); o.rowDx.push_back(o.angleIdx[3*t+b]);
        }
    for (int t = 0; t < o.nw; t++)
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
            o.rowDq.push_back(o.waterIdx[3*t+a]); o.rowDx.push_back(o.waterIdx[3*t+b]);
        }
    o.rowVal.assign(3*o.rowDq.size(), 0.0);
    o.excl.resize(n);
    for (int e = 0; e < d->num_exceptions; e++) {
        int a = d->exception_pairs[2*e], b = d->exception_pairs[2*e+1];
        if (a < 0 || a >= n || b < 0 || b >= n) { delete h; return fail(CFX_ERR_ARGUMENT, "exception index out of range"); }
        o.excl[a].insert(b);
        o.excl[b].insert(a);
    }
    o.pbc = d->use_pbc != 0;
    if (o.pbc) {
        o.cutoff = d->cutoff;
        o.tol = d->ewald_tol;
        o.alpha = (1.0/o.cutoff)*sqrt(-log(2.0*o.tol));               // :401
        for (int a = 0; a < 3; a++) {                                   // :403-420
            int k = 1;
            while (ewaldErrorEstimate(k, d->default_box[4*a], o.alpha) > o.tol)
                k++;
            if (k%2 == 0) k++;
            o.kmax[a] = k;
        }
    }
    o.q.assign(n, 0.0); o.dedq.assign(n, 0.0);
    *out = h;
    return CFX_OK;
}

void cfxo_destroy(cfxo_handle* h) { delete h; }

/* Restrict the reciprocal loop to nkx in [lo,hi) -- used only to time a bounded sample of the CPU
 * baseline (bench.py). hi < 0 restores the full range. */
int cfxo_set_kx_range(cfxo_handle* h, int lo, int hi) { h->o.kxLo = lo; h->o.kxHi = hi; return CFX_OK; }

/* ReferenceCoulKernels.cpp:424-636. forces (may be NULL) is ADDED to. */
int cfxo_execute(cfxo_handle* h, const double* positions, const double* box, int includeForces, int includeEnergy,
                 double* energy, double* forces) {
    Oracle& o = h->o;
    const V3* pos = reinterpret_cast<const V3*>(positions);
    std::vector<V3> scratch;
    V3* F = reinterpret_cast<V3*>(forces);
    if (!F) { scratch.assign(o.n, V3{0,0,0}); F = scratch.data(); }
    double e5[CFX_E_COUNT] = {0,0,0,0,0};
    assembleCharges(o, pos, box);
    o.dedq.assign(o.n, 0.0);
    if (o.pbc) {
        if (box[1] != 0 || box[2] != 0 || box[3] != 0 || box[5] != 0 || box[6] != 0 || box[7] != 0)
            return fail(CFX_ERR_ARGUMENT, "only rectangular boxes are supported");
        evaluateEwald(o, pos, box, includeForces != 0, includeEnergy != 0, e5, F);
    }
    else
        evaluateNoCutoff(o, pos, includeForces != 0, includeEnergy != 0, e5, F);
    // chain rule, :493-499 / :626-632 -- always applied, with whatever dE/dq was accumulated.
    for (size_t r = 0; r < o.rowDq.size(); r++) {
        int a = o.rowDq[r], b = o.rowDx[r];
        for (int c = 0; c < 3; c++)
            F[b][c] -= o.dedq[a]*o.rowVal[3*r+c];
    }
    if (energy) memcpy(energy, e5, sizeof(e5));
    return CFX_OK;
}

int cfxo_get_ewald_params(const cfxo_handle* h, cfx_ewald_params* out) {
    const Oracle& o = h->o;
    out->alpha = o.alpha;
    for (int a = 0; a < 3; a++) out->kmax[a] = o.kmax[a];
    long long kx = o.kmax[0], ky = o.kmax[1], kz = o.kmax[2];
    out->num_kvectors = o.pbc ? (kz - 1) + (ky - 1)*(2*kz - 1) + (kx - 1)*(2*ky - 1)*(2*kz - 1) : 0;
    return CFX_OK;
}

int cfxo_get_charges(cfxo_handle* h, double* q) { memcpy(q, h->o.q.data(), sizeof(double)*h->o.n); return CFX_OK; }
int cfxo_get_dedq(cfxo_handle* h, double* v) { memcpy(v, h->o.dedq.data(), sizeof(double)*h->o.n); return CFX_OK; }
int cfxo_num_jacobian_rows(const cfxo_handle* h) { return (int) h->o.rowDq.size(); }
int cfxo_get_jacobian(cfxo_handle* h, int32_t* dq, int32_t* dx, double* val) {
    const Oracle& o = h->o;
    if (dq) memcpy(dq, o.rowDq.data(), sizeof(int)*o.rowDq.size());
    if (dx) memcpy(dx, o.rowDx.data(), sizeof(int)*o.rowDx.size());
    if (val) memcpy(val, o.rowVal.data(), sizeof(double)*o.rowVal.size());
    return CFX_OK;
}
int cfxo_get_neighbor_pairs(cfxo_handle* h, int32_t* pairs, int64_t capacity, int64_t* count) {
    Oracle& o = h->o;
    *count = (int64_t) o.pairs.size();
    if (!pairs) return CFX_OK;
    if (capacity < *count) return fail(CFX_ERR_ARGUMENT, "pair buffer too small");
    std::vector<std::pair<int,int> > sorted(o.pairs);
    std::sort(sorted.begin(), sorted.end());
    for (size_t k = 0; k < sorted.size(); k++) { pairs[2*k] = sorted[k].first; pairs[2*k+1] = sorted[k].second; }
    return CFX_OK;
}
int cfxo_get_stats(const cfxo_handle* h, cfx_stats* out) {
    memset(out, 0, sizeof(*out));
    out->pairs_in_cutoff = (int64_t) h->o.pairs.size();
    out->pair_candidates = h->o.pairCandidates;
    return CFX_OK;
}

} // extern "C"
