#include "model.h"

#include <cmath>
#include <cstring>
#include <mutex>

namespace pml {

namespace {

// WAG (Whelan & Goldman 2001) exchangeabilities, PAML wag.dat scale x100, lower triangle,
// residue order A R N D C Q E G H I L K M F P S T W Y V.
const double kWagLower[190] = {
    55.1571,  50.9848, 63.5346, 73.8998, 14.7304, 542.942, 102.704, 52.8191, 26.5256, 3.02949, 90.8598, 303.55,
    154.364,  61.6783, 9.88179, 158.285, 43.9157, 94.7198, 617.416, 2.1352,  546.947, 141.672, 58.4665, 112.556,
    86.5584,  30.6674, 33.0052, 56.7717, 31.6954, 213.715, 395.629, 93.0676, 24.8972, 429.411, 57.0025, 24.941,
    19.3335,  18.6979, 55.4236, 3.9437,  17.0135, 11.3917, 12.7395, 3.04501, 13.819,  39.7915, 49.7671, 13.1528,
    8.48047,  38.4287, 86.9489, 15.4263, 6.13037, 49.9462, 317.097, 90.6265, 535.142, 301.201, 47.9855, 7.40339,
    389.49,   258.443, 37.3558, 89.0432, 32.3832, 25.7555, 89.3496, 68.3162, 19.8221, 10.3754, 39.0482, 154.526,
    31.5124,  17.41,   40.4141, 425.746, 485.402, 93.4276, 21.0494, 10.2711, 9.61621, 4.67304, 39.802,  9.99208,
    8.11339,  4.9931,  67.9371, 105.947, 211.517, 8.8836,  119.063, 143.855, 67.9489, 19.5081, 42.3984, 10.9404,
    93.3372,  68.2355, 24.357,  69.6198, 9.99288, 41.5844, 55.6896, 17.1329, 16.1444, 337.079, 122.419, 397.423,
    107.176,  140.766, 102.887, 70.4939, 134.182, 74.0169, 31.944,  34.4739, 96.713,  49.3905, 54.5931, 161.328,
    212.111,  55.4413, 203.006, 37.4866, 51.2984, 85.7928, 82.2765, 22.5833, 47.3307, 145.816, 32.6622, 138.698,
    151.612,  17.1903, 79.5384, 437.802, 11.3133, 116.392, 7.19167, 12.9767, 71.707,  21.5737, 15.6557, 33.6983,
    26.2569,  21.2483, 66.5309, 13.7505, 51.5706, 152.964, 13.9405, 52.3742, 11.0864, 24.0735, 38.1533, 108.6,
    32.5711,  54.3833, 22.771,  19.6303, 10.3604, 387.344, 42.017,  39.8618, 13.3264, 42.8437, 645.428, 21.6046,
    78.6993,  29.1148, 248.539, 200.601, 25.1849, 19.6246, 15.2335, 100.214, 30.1281, 58.8731, 18.7247, 11.8358,
    782.13,   180.034, 30.5434, 205.845, 64.9892, 31.4887, 23.2739, 138.823, 36.5369, 31.473};
const double kWagPi[kStates] = {0.087, 0.044, 0.039, 0.057, 0.019, 0.037, 0.058, 0.083, 0.024, 0.049,
                                0.086, 0.062, 0.020, 0.038, 0.046, 0.070, 0.061, 0.014, 0.035, 0.071};

// Symmetric eigenproblem by one-sided threshold Jacobi sweeps on M (destroyed); vectors accumulate in R's columns.
void symmetric_eigen(double M[kStates][kStates], double R[kStates][kStates], double* w) {
    const int n = kStates;
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) R[a][b] = a == b ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double rest = 0.0;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) rest += std::fabs(M[p][q]);
        if (rest == 0.0) break;
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                const double apq = M[p][q];
                if (apq == 0.0) continue;
                const double diff = M[q][q] - M[p][p];
                double t;
                if (std::fabs(apq) < 1e-40 * std::fabs(diff)) {
                    t = apq / diff;
                } else {
                    const double phi = diff / (2.0 * apq);
                    t = 1.0 / (std::fabs(phi) + std::sqrt(phi * phi + 1.0));
                    if (phi < 0.0) t = -t;
                }
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c, tau = s / (1.0 + c);
                M[p][p] -= t * apq;
                M[q][q] += t * apq;
                M[p][q] = M[q][p] = 0.0;
                for (int r = 0; r < n; ++r) {
                    if (r != p && r != q) {
                        const double mrp = M[r][p], mrq = M[r][q];
                        M[r][p] = M[p][r] = mrp - s * (mrq + tau * mrp);
                        M[r][q] = M[q][r] = mrq + s * (mrp - tau * mrq);
                    }
                    const double vrp = R[r][p], vrq = R[r][q];
                    R[r][p] = vrp - s * (vrq + tau * vrp);
                    R[r][q] = vrq + s * (vrp - tau * vrq);
                }
            }
        }
    }
    for (int a = 0; a < n; ++a) w[a] = M[a][a];
}

Eigensystem build_wag() {
    Eigensystem es{};
    double S[kStates][kStates] = {};
    int k = 0;
    for (int r = 1; r < kStates; ++r)
        for (int c = 0; c < r; ++c) S[r][c] = S[c][r] = kWagLower[k++];
    std::memcpy(es.pi, kWagPi, sizeof kWagPi);
    // Q_ij = S_ij pi_j, rows sum to zero, scaled so that -sum_i pi_i Q_ii = 1
    double Q[kStates][kStates];
    double flux = 0.0;
    for (int i = 0; i < kStates; ++i) {
        double out = 0.0;
        for (int j = 0; j < kStates; ++j) {
            Q[i][j] = i == j ? 0.0 : S[i][j] * es.pi[j];
            out += Q[i][j];
        }
        Q[i][i] = -out;
        flux += es.pi[i] * out;
    }
    // similarity transform with sqrt(pi) makes Q symmetric: B = D^{1/2} Q D^{-1/2}
    double B[kStates][kStates], R[kStates][kStates], root[kStates];
    for (int i = 0; i < kStates; ++i) root[i] = std::sqrt(es.pi[i]);
    for (int i = 0; i < kStates; ++i)
        for (int j = 0; j <= i; ++j) {
            const double v = i == j ? Q[i][i] / flux : S[i][j] * root[i] * root[j] / flux;
            B[i][j] = B[j][i] = v;
        }
    symmetric_eigen(B, R, es.lambda);
    for (int i = 0; i < kStates; ++i)
        for (int e = 0; e < kStates; ++e) {
            es.V[i][e] = R[i][e] / root[i];
            es.Vinv[e][i] = R[i][e] * root[i];
        }
    for (int e = 0; e < kStates; ++e)
        if (es.lambda[e] > 0.0) es.lambda[e] = 0.0;  // the stationary mode; rounding may leave +1e-17
    return es;
}

// regularised lower incomplete gamma P(a, x): power series below a+1, modified Lentz continued fraction above
double gamma_p(double a, double x) {
    if (!(x > 0.0)) return 0.0;
    const double front = std::exp(a * std::log(x) - x - std::lgamma(a));
    if (x < a + 1.0) {
        double term = 1.0 / a, total = term;
        for (int n = 1; n < 200000; ++n) {
            term *= x / (a + n);
            total += term;
            if (term < total * 1e-18) break;
        }
        return front * total;
    }
    const double tiny = 1e-290;
    double b = x + 1.0 - a, C = 1.0 / tiny, D = 1.0 / b, f = D;
    for (int n = 1; n < 200000; ++n) {
        const double an = -n * (n - a);
        b += 2.0;
        D = an * D + b;
        if (std::fabs(D) < tiny) D = tiny;
        C = b + an / C;
        if (std::fabs(C) < tiny) C = tiny;
        D = 1.0 / D;
        const double step = D * C;
        f *= step;
        if (std::fabs(step - 1.0) < 1e-17) break;
    }
    return 1.0 - front * f;
}

// x with P(a, x) = p : safeguarded Newton (Halley-free) inside a shrinking bracket
double gamma_p_inverse(double a, double p) {
    double lo = 0.0, hi = a > 1.0 ? a : 1.0;
    while (gamma_p(a, hi) < p) {
        lo = hi;
        hi *= 2.0;
    }
    double x = 0.5 * (lo + hi);
    const double lg = std::lgamma(a);
    for (int it = 0; it < 300; ++it) {
        const double err = gamma_p(a, x) - p;
        if (err > 0.0) hi = x; else lo = x;
        const double dens = std::exp((a - 1.0) * std::log(x) - x - lg);
        double next = dens > 0.0 ? x - err / dens : 0.5 * (lo + hi);
        if (!(next > lo && next < hi)) next = 0.5 * (lo + hi);
        if (std::fabs(next - x) <= 4e-16 * x || hi - lo <= 4e-16 * hi) {
            x = next;
            break;
        }
        x = next;
    }
    return x;
}

}  // namespace

const Eigensystem& wag_eigensystem() {
    static const Eigensystem es = build_wag();
    return es;
}

void wag_pmatrix(double t, double rate, double* P) {
    const Eigensystem& es = wag_eigensystem();
    double scaled[kStates][kStates];
    for (int e = 0; e < kStates; ++e) {
        const double g = std::exp(es.lambda[e] * rate * t);
        for (int j = 0; j < kStates; ++j) scaled[e][j] = g * es.Vinv[e][j];
    }
    for (int i = 0; i < kStates; ++i)
        for (int j = 0; j < kStates; ++j) {
            double acc = 0.0;
            for (int e = 0; e < kStates; ++e) acc += es.V[i][e] * scaled[e][j];
            P[i * kStates + j] = acc;
        }
}

// Yang (1994) discrete Gamma with the MEAN of each equal-probability bin: the reference uses means, not medians
// (median rates miss the oracle lnL by 0.03 on the 8x300 case).
void gamma_mean_rates(double alpha, int ncat, double* rates) {
    double below = 0.0;
    for (int k = 0; k < ncat; ++k) {
        double upto = 1.0;
        if (k + 1 < ncat) {
            // quantile of Gamma(shape alpha, rate alpha) is q/alpha with q the unit-rate quantile; the mass of
            // x*f(x) below it is P(alpha + 1, q)
            const double q = gamma_p_inverse(alpha, double(k + 1) / ncat);
            upto = gamma_p(alpha + 1.0, q);
        }
        rates[k] = (upto - below) * ncat;
        below = upto;
    }
}

int residue_code(unsigned char ch) {
    static const std::array<signed char, 256> table = [] {
        std::array<signed char, 256> t{};
        t.fill(22);
        const char* order = "ARNDCQEGHILKMFPSTWYV";
        for (int i = 0; i < kStates; ++i) {
            t[(unsigned char)order[i]] = (signed char)i;
            t[(unsigned char)(order[i] + 32)] = (signed char)i;
        }
        t['B'] = t['b'] = 20;
        t['Z'] = t['z'] = 21;
        return t;
    }();
    return table[ch];
}

void code_indicator(int code, double* v) {
    for (int j = 0; j < kStates; ++j) v[j] = code >= 22 ? 1.0 : 0.0;
    if (code < 20) v[code] = 1.0;
    else if (code == 20) v[2] = v[3] = 1.0;   // B = N or D
    else if (code == 21) v[5] = v[6] = 1.0;   // Z = Q or E
}

}  // namespace pml
