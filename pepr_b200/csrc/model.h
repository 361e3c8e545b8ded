// Host-side substitution model for the engine: WAG exchangeabilities with the fixed 3-decimal WAG frequencies,
// Q normalised to one expected substitution per site, symmetric eigendecomposition, mean discrete-Gamma rates.
// Takes over raxmlHPC initProtMat/putWAG/initReversibleGTR/makeGammaCats.
#pragma once
#include <array>
#include <string>
#include <vector>

namespace pml {

constexpr int kStates = 20;
constexpr int kCats = 4;
constexpr int kRow = kStates * kCats;  // one CLV row: 4 categories x 20 states
constexpr int kCodes = 23;             // 20 residues, B, Z, undetermined

struct Eigensystem {
    double pi[kStates];
    double lambda[kStates];           // eigenvalues of Q, all <= 0
    double V[kStates][kStates];       // right eigenvectors in columns: Q = V diag(lambda) Vinv
    double Vinv[kStates][kStates];
};

const Eigensystem& wag_eigensystem();               // computed once, thread safe
void wag_pmatrix(double t, double rate, double* P);  // row-major P(i->j)
void gamma_mean_rates(double alpha, int ncat, double* rates);
int residue_code(unsigned char ch);                 // 0..19 residues, 20 = B, 21 = Z, 22 = undetermined
// indicator vector of a tip code (1 for each residue the code allows)
void code_indicator(int code, double* v20);

constexpr double kAlphaMin = 0.02, kAlphaMax = 1000.0;

}  // namespace pml
