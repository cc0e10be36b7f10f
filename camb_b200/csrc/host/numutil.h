// numutil.h -- host-side numerical utilities of the stand-in host:
// Romberg integration, Brent root finding, cubic splines and the TRanges
// piecewise-uniform grids.  Each routine states the reference routine whose
// behaviour it reproduces; none of this is on the GPU hot path.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <limits>
#include <stdexcept>
#include <vector>

namespace cbh {

using fn1 = std::function<double(double)>;

inline int f_nint(double x) { return (int)std::lround(x); }  // Fortran nint: half away from zero

// fortran/MathUtils.f90:648-723 Integrate_Romberg (relative tolerance unless abs_tol)
inline double romberg(const fn1& f, double a, double b, double tol, int max_it = 25, int min_steps = 0,
                      bool abstol = false) {
    const int MAXJ = 5;
    double g[MAXJ + 2];
    double h = 0.5 * (b - a);
    double gmax = h * (f(a) + f(b));
    g[1] = gmax;
    double g0 = gmax;
    long nint = 1;
    double error = 1.0e20;
    int i = 0;
    for (;;) {
        i++;
        if (i > max_it || ((i > 5 && std::fabs(error) < tol) && nint > min_steps)) break;
        g0 = 0.0;
        for (long k = 1; k <= nint; k++) g0 += f(a + (double)(k + k - 1) * h);
        g0 = 0.5 * g[1] + h * g0;
        h = 0.5 * h;
        nint = nint + nint;
        int jmax = i < MAXJ ? i : MAXJ;
        double fourj = 1.0;
        for (int j = 1; j <= jmax; j++) {
            fourj = 4.0 * fourj;
            double g1 = g0 + (g0 - g[j]) / (fourj - 1.0);
            g[j] = g0;
            g0 = g1;
        }
        if (abstol) error = std::fabs(gmax - g0);
        else if (std::fabs(g0) > tol) error = 1.0 - gmax / g0;
        else error = gmax;
        gmax = g0;
        g[jmax + 1] = g0;
    }
    return g0;
}

// fortran/MathUtils.f90:725-880 brentq (zeroin).  Returns false if no sign change.
inline bool brentq(const fn1& f, double ax, double bx, double tol, double& xzero, double& fzero,
                   const double* fax = nullptr, const double* fbx = nullptr) {
    const double eps = std::numeric_limits<double>::epsilon();
    double a = ax, b = bx;
    double fa = fax ? *fax : f(a);
    double fb = fbx ? *fbx : f(b);
    if (fa == 0) { xzero = a; fzero = fa; return true; }
    if (fb == 0) { xzero = b; fzero = fb; return true; }
    if (!(fa * (fb / std::fabs(fb)) < 0)) return false;
    double c = a, fc = fa, d = b - a, e = d;
    for (;;) {
        if (std::fabs(fc) < std::fabs(fb)) { a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
        double tol1 = 2 * eps * std::fabs(b) + 0.5 * tol;
        double xm = 0.5 * (c - b);
        if (std::fabs(xm) <= tol1 || fb == 0) break;
        if (std::fabs(e) >= tol1 && std::fabs(fa) > std::fabs(fb)) {
            double s = fb / fa, p, q;
            if (a != c) {
                q = fa / fc;
                double r = fb / fc;
                p = s * (2 * xm * q * (q - r) - (b - a) * (r - 1));
                q = (q - 1) * (r - 1) * (s - 1);
            } else { p = 2 * xm * s; q = 1 - s; }
            if (p <= 0) p = -p; else q = -q;
            s = e; e = d;
            if ((2 * p) >= (3 * xm * q - std::fabs(tol1 * q)) || p >= std::fabs(0.5 * s * q)) { d = xm; e = d; }
            else d = p / q;
        } else { d = xm; e = d; }
        a = b; fa = fb;
        if (std::fabs(d) <= tol1) b += (xm <= 0 ? -tol1 : tol1); else b += d;
        fb = f(b);
        if (fb * (fc / std::fabs(fc)) > 0) { c = a; fc = fa; d = b - a; e = d; }
    }
    xzero = b; fzero = fb;
    return true;
}

constexpr double SPLINE_DANGLE = 1.0e30;

// forutils/Interpolation.f90:919-964 spline (natural end if d11/d1n == SPLINE_DANGLE)
inline void spline(const double* x, const double* y, int n, double d11, double d1n, double* d2) {
    std::vector<double> u(n > 1 ? n - 1 : 1);
    double d1r = (y[1] - y[0]) / (x[1] - x[0]), d1l;
    if (d11 == SPLINE_DANGLE) { d2[0] = 0; u[0] = 0; }
    else { d2[0] = -0.5; u[0] = (3.0 / (x[1] - x[0])) * (d1r - d11); }
    for (int i = 1; i < n - 1; i++) {
        d1l = d1r;
        d1r = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
        double xxdiv = 1.0 / (x[i + 1] - x[i - 1]);
        double sig = (x[i] - x[i - 1]) * xxdiv;
        double xp = 1.0 / (sig * d2[i - 1] + 2.0);
        d2[i] = (sig - 1.0) * xp;
        u[i] = (6.0 * (d1r - d1l) * xxdiv - sig * u[i - 1]) * xp;
    }
    d1l = d1r;
    double qn, un;
    if (d1n == SPLINE_DANGLE) { qn = 0; un = 0; }
    else { qn = 0.5; un = (3.0 / (x[n - 1] - x[n - 2])) * (d1n - d1l); }
    d2[n - 1] = (un - qn * u[n - 2]) / (qn * d2[n - 2] + 1.0);
    for (int i = n - 2; i >= 0; i--) d2[i] = d2[i] * d2[i + 1] + u[i];
}

// forutils/Interpolation.f90:1072-1116 regular_spline
inline void regular_spline(double delta, const double* y, int n, double d11, double d1n, double* d2) {
    std::vector<double> u(n - 1);
    double inv_delta = 1.0 / delta, three_inv_delta = 3.0 * inv_delta;
    double d1r = (y[1] - y[0]) * inv_delta, d1l;
    if (d11 == SPLINE_DANGLE) { d2[0] = 0; u[0] = 0; }
    else { d2[0] = -0.5; u[0] = three_inv_delta * (d1r - d11); }
    for (int i = 1; i < n - 1; i++) {
        d1l = d1r;
        d1r = (y[i + 1] - y[i]) * inv_delta;
        double xp = 1.0 / (0.5 * d2[i - 1] + 2.0);
        d2[i] = -0.5 * xp;
        u[i] = (three_inv_delta * (d1r - d1l) - 0.5 * u[i - 1]) * xp;
    }
    d1l = d1r;
    double qn, un;
    if (d1n == SPLINE_DANGLE) { qn = 0; un = 0; }
    else { qn = 0.5; un = three_inv_delta * (d1n - d1l); }
    d2[n - 1] = (un - qn * u[n - 2]) / (qn * d2[n - 2] + 1.0);
    for (int i = n - 2; i >= 0; i--) d2[i] = d2[i] * d2[i + 1] + u[i];
}

// forutils/Interpolation.f90:1053-1070 (4 x (n-1), column-major -> horner[4*i+c])
inline void regular_horner(double delta, const double* y, const double* d2, double* horner, int n) {
    double h2over6 = delta * delta / 6.0, three = 3.0 * h2over6;
    for (int i = 0; i < n - 1; i++) {
        horner[4 * i + 0] = y[i];
        horner[4 * i + 1] = y[i + 1] - y[i] - h2over6 * (2.0 * d2[i] + d2[i + 1]);
        horner[4 * i + 2] = three * d2[i];
        horner[4 * i + 3] = h2over6 * (d2[i + 1] - d2[i]);
    }
}

// TCubicSpline subset (forutils/Interpolation.f90:274-365, 418-462, 519-532)
struct CubicSpline {
    std::vector<double> X, F, ddF;
    void init(const std::vector<double>& x, const std::vector<double>& f) {
        X = x; F = f; ddF.assign(x.size(), 0.0);
        spline(X.data(), F.data(), (int)X.size(), SPLINE_DANGLE, SPLINE_DANGLE, ddF.data());
    }
    // IntegralArray(yint, first_index): yint[0]=0 corresponds to X[first]
    void integral_array(double* yint, int first = 0) const {
        yint[0] = 0;
        int ix = 1;
        for (size_t i = first + 1; i < X.size(); i++, ix++) {
            double dx = X[i] - X[i - 1];
            yint[ix] = yint[ix - 1] + dx * ((F[i] + F[i - 1]) / 2 - dx * dx / 24 * (ddF[i] + ddF[i - 1]));
        }
    }
    // ArrayValue for monotonically increasing x (TSpline1D_ArrayValue + TCubicSpline_FindNext)
    void array(const double* x, double* y, int n) const {
        size_t llo = 0;
        for (int i = 0; i < n; i++) {
            while (llo + 1 < X.size() - 1 && X[llo + 1] < x[i]) llo++;
            double xlo = X[llo], xhi = X[llo + 1];
            double ho = xhi - xlo, a0 = (xhi - x[i]) / ho, b0 = 1 - a0;
            y[i] = b0 * F[llo + 1] + a0 * (F[llo] - b0 * ((a0 + 1) * ddF[llo] + (2 - a0) * ddF[llo + 1]) * ho * ho / 6);
        }
    }
};

// forutils/RangeUtils.f90 TRange / TRanges
struct TRange {
    int start_index = 0, steps = 0;
    bool IsLog = false;
    double Low = 0, High = 0, delta = 0, delta_max = 0, delta_min = 0;
};

struct TRanges {
    int count = 0, npoints = 0;
    double Lowest = 0, Highest = 0;
    std::vector<TRange> R;
    std::vector<double> points, dpoints;
    double RangeTol = 0.1;

    void init() { R.clear(); points.clear(); dpoints.clear(); count = 0; npoints = 0; }

    static void setmins(TRange& r) {
        if (r.IsLog) {
            if (r.steps == 1) { r.delta_min = r.High - r.Low; r.delta_max = r.delta_min; }
            else { r.delta_min = r.Low * (std::exp(r.delta) - 1); r.delta_max = r.High * (1 - std::exp(-r.delta)); }
        } else { r.delta_max = r.delta; r.delta_min = r.delta; }
    }
    void set_delta(TRange& r, double delta) const {
        double Diff = r.IsLog ? std::log(r.High / r.Low) : r.High - r.Low;
        if (delta >= Diff) { r.delta = Diff; r.steps = 1; }
        else {
            int s = (int)(Diff / delta + 1.0 - RangeTol);
            r.steps = s > 1 ? s : 1;
            r.delta = Diff / r.steps;
        }
        setmins(r);
    }

    // RangeUtils.f90:171-191
    void add_delta(double t_start, double t_end, double t_approx_delta, bool IsLog = false) {
        if (t_end <= t_start) throw std::runtime_error("TRanges_Add_delta: end must be larger than start");
        if (t_approx_delta <= 0) throw std::runtime_error("TRanges_Add_delta: delta must be > 0");
        int n;
        if (IsLog) n = (int)(std::log(t_end / t_start) / t_approx_delta + 1.0 - RangeTol);
        else n = (int)((t_end - t_start) / t_approx_delta + 1.0 - RangeTol);
        if (n < 1) n = 1;
        add(t_start, t_end, n, IsLog);
    }

    // RangeUtils.f90:193-537
    void add(double t_start, double t_end, int nstep, bool WantLog = false) {
        if (t_end <= t_start) throw std::runtime_error("TRanges_Add: end must be larger than start");
        if (nstep <= 0) throw std::runtime_error("TRanges_Add: nstep must be > 0");
        double delta = WantLog ? std::log(t_end / t_start) / nstep : (t_end - t_start) / nstep;
        std::vector<TRange> NewRanges;
        TRange NewR;
        NewR.Low = t_start; NewR.High = t_end; NewR.delta = delta; NewR.steps = nstep; NewR.IsLog = WantLog;
        setmins(NewR);
        for (int i = 0; i < count; i++) {
            TRange& Rr = R[i];
            if (NewR.Low >= Rr.Low && NewR.High <= Rr.High &&
                (NewR.delta_min >= Rr.delta_max || ((NewR.IsLog == Rr.IsLog) && NewR.delta >= Rr.delta))) {
                return;  // new range narrower than existing and with wider spacing
            } else if (NewR.Low <= Rr.Low && NewR.High >= Rr.High &&
                       (NewR.delta_min <= Rr.delta_min || ((NewR.IsLog == Rr.IsLog) && NewR.delta <= Rr.delta))) {
                continue;  // R completely inside and wider than new one
            } else if (NewR.delta_min < Rr.delta_max && NewR.High > Rr.Low && NewR.Low < Rr.High) {
                if (Rr.IsLog && !NewR.IsLog) {
                    int nsteps = (int)(std::log(NewR.delta / (std::exp(Rr.delta) - 1) / Rr.Low) / Rr.delta + 1);
                    double splitpt = Rr.Low * std::exp(Rr.delta * nsteps);
                    if (splitpt * (1 - std::exp(-Rr.delta)) > NewR.delta * (1 + RangeTol))
                        splitpt = NewR.delta / (1 - std::exp(-Rr.delta));
                    if (splitpt < Rr.High) {
                        if (Rr.High > NewR.High) {
                            TRange T; T.Low = NewR.High; T.High = Rr.High; T.IsLog = Rr.IsLog;
                            set_delta(T, Rr.delta);
                            NewRanges.push_back(T);
                        }
                        Rr.High = std::max(NewR.Low, splitpt);
                        set_delta(Rr, Rr.delta);
                    }
                } else if (!Rr.IsLog && NewR.IsLog) {
                    int nsteps = (int)(std::log(Rr.delta / (std::exp(NewR.delta) - 1) / NewR.Low) / NewR.delta + 1);
                    double splitpt = NewR.Low * std::exp(NewR.delta * nsteps);
                    if (splitpt * (1 - std::exp(-NewR.delta)) > Rr.delta * (1 + RangeTol))
                        splitpt = Rr.delta / (1 - std::exp(-NewR.delta));
                    if (splitpt < Rr.High && splitpt > Rr.Low) {
                        if (Rr.Low < NewR.Low) {
                            TRange T; T.Low = Rr.Low; T.High = NewR.Low; T.IsLog = Rr.IsLog;
                            set_delta(T, Rr.delta);
                            NewRanges.push_back(T);
                        }
                        Rr.Low = std::min(NewR.High, splitpt);
                        set_delta(Rr, Rr.delta);
                    }
                }
            }
            NewRanges.push_back(Rr);
        }
        NewRanges.push_back(NewR);
        int nreg = (int)NewRanges.size();

        std::vector<double> EndPoints(2 * nreg + 3, 0.0);  // 1-based use
        int ix = 0;
        for (int i = 0; i < nreg; i++) {
            const TRange& A = NewRanges[i];
            if (ix == 0) { EndPoints[1] = A.Low; EndPoints[2] = A.High; ix = 2; }
            else {
                int ixin = ix;
                for (int j = 1; j <= ixin; j++) {
                    if (A.Low < EndPoints[j]) {
                        for (int k = ix + 1; k >= j + 1; k--) EndPoints[k] = EndPoints[k - 1];
                        EndPoints[j] = A.Low; ix++;
                        break;
                    }
                }
                if (ixin == ix) {
                    ix++; EndPoints[ix] = A.Low;
                    ix++; EndPoints[ix] = A.High;
                } else {
                    ixin = ix;
                    for (int j = 1; j <= ixin; j++) {
                        if (A.High < EndPoints[j]) {
                            for (int k = ix + 1; k >= j + 1; k--) EndPoints[k] = EndPoints[k - 1];
                            EndPoints[j] = A.High; ix++;
                            break;
                        }
                    }
                    if (ixin == ix) { ix++; EndPoints[ix] = A.High; }
                }
            }
        }
        int ixin = ix;
        ix = 1;
        for (int i = 2; i <= ixin; i++)
            if (EndPoints[i] != EndPoints[ix]) { ix++; EndPoints[ix] = EndPoints[i]; }

        Lowest = EndPoints[1];
        Highest = EndPoints[ix];
        count = 0;
        double max_delta = Highest - Lowest;
        R.assign(ix, TRange());
        std::vector<double> RequestDelta(ix + 1, 0.0);
        for (int i = 1; i <= ix - 1; i++) {
            TRange& A = R[i - 1];
            A.Low = EndPoints[i]; A.High = EndPoints[i + 1];
            double dl = max_delta;
            A.IsLog = false;
            for (int j = 0; j < nreg; j++) {
                const TRange& N = NewRanges[j];
                if (A.Low >= N.Low && A.Low < N.High) {
                    if (N.IsLog) {
                        if (A.IsLog) dl = std::min(dl, N.delta);
                        else {
                            double min_log_step = A.Low * (std::exp(N.delta) - 1);
                            if (min_log_step < dl) {
                                double max_log_step = A.High * (1 - std::exp(-N.delta));
                                A.IsLog = true;
                                if (max_log_step <= dl) dl = N.delta;
                                else dl = -std::log(1 - dl / A.High);
                            }
                        }
                    } else {
                        if (A.IsLog) {
                            double max_log_step = A.High * (1 - std::exp(-dl));
                            if (N.delta < max_log_step) {
                                double min_log_step = A.Low * (std::exp(dl) - 1);
                                if (min_log_step < N.delta) dl = -std::log(1 - N.delta / A.High);
                                else { A.IsLog = false; dl = N.delta; }
                            }
                        } else dl = std::min(dl, N.delta);
                    }
                }
            }
            set_delta(A, dl);
            count++;
            RequestDelta[count] = dl;
        }

        // Get rid of tiny ranges (RangeUtils.f90:403-457); indices 1-based on R[i-1]
        int ixc = count;
        for (int i = ixc; i >= 1; i--) {
            TRange& A = R[i - 1];
            if (A.steps == 1) {
                double diff = A.High - A.Low, min_request, max_request;
                if (A.IsLog) {
                    min_request = A.Low * (std::exp(RequestDelta[i]) - 1);
                    max_request = A.High * (1 - std::exp(-RequestDelta[i]));
                } else { min_request = RequestDelta[i]; max_request = min_request; }
                if (i != count) {
                    TRange& L = R[i];
                    if (RequestDelta[i] >= A.delta && diff <= L.delta_min && L.delta_min <= max_request &&
                        (!L.IsLog || A.Low > L.Low * std::exp(-L.delta))) {
                        L.Low = A.Low;
                        if (diff > L.delta_min * RangeTol) L.steps++;
                        if (L.IsLog) L.delta = std::log(L.High / L.Low) / L.steps;
                        else L.delta = (L.High - L.Low) / L.steps;
                        for (int k = i; k <= count - 1; k++) R[k - 1] = R[k];
                        count--;
                        continue;
                    }
                }
                if (i != 1) {
                    TRange& L = R[i - 2];
                    if (RequestDelta[i] >= A.delta && diff <= L.delta_max &&
                        (L.delta_max <= min_request || diff < L.delta_min * RangeTol)) {
                        L.High = A.High;
                        if (diff > L.delta_max * RangeTol) L.steps++;
                        if (L.IsLog) L.delta = std::log(L.High / L.Low) / L.steps;
                        else L.delta = (L.High - L.Low) / L.steps;
                        for (int k = i; k <= count - 1; k++) R[k - 1] = R[k];
                        count--;
                    }
                }
            }
        }
        int nsteps = 1;
        for (int i = 0; i < count; i++) {
            TRange& A = R[i];
            A.start_index = nsteps;
            nsteps += A.steps;
            setmins(A);
        }
        R.resize(count);
        npoints = nsteps;
    }

    // RangeUtils.f90:119-169
    void get_array(bool want_dpoints = true, bool half_ends = true) {
        points.assign(npoints, 0.0);
        int ix = 0;
        for (int i = 0; i < count; i++) {
            const TRange& A = R[i];
            for (int j = 0; j < A.steps; j++) {
                points[ix++] = A.IsLog ? A.Low * std::exp(j * A.delta) : A.Low + A.delta * j;
            }
        }
        points[ix++] = Highest;
        if (ix != npoints) throw std::runtime_error("TRanges_GetArray: ERROR");
        if (want_dpoints) get_dpoints(half_ends);
    }
    void get_dpoints(bool half_ends = true) {
        dpoints.assign(npoints, 0.0);
        for (int i = 1; i < npoints - 1; i++) dpoints[i] = (points[i + 1] - points[i - 1]) / 2;
        if (half_ends) {
            dpoints[0] = (points[1] - points[0]) / 2;
            dpoints[npoints - 1] = (points[npoints - 1] - points[npoints - 2]) / 2;
        } else {
            dpoints[0] = points[1] - points[0];
            dpoints[npoints - 1] = points[npoints - 1] - points[npoints - 2];
        }
    }
    // RangeUtils.f90:68-94, returns 1-based index
    int index_of(double tau) const {
        for (int i = 0; i < count; i++) {
            const TRange& A = R[i];
            if (tau < A.High && tau >= A.Low) {
                if (A.IsLog) return A.start_index + (int)(std::log(tau / A.Low) / A.delta);
                return A.start_index + (int)((tau - A.Low) / A.delta);
            }
        }
        if (tau >= Highest) return npoints;
        throw std::runtime_error("TRanges_IndexOf: value out of range");
    }
};

}  // namespace cbh
