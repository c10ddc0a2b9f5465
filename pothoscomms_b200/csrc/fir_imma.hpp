// Bit-exact int16 / complex-int16 FIR on the int8 tensor cores (fir_imma.cu).
#pragma once
#include <cstddef>
#include <cstdint>

#include "common.hpp"

namespace b200c {

// The reference accumulates int16 data times Q16 taps in wrapping int32 (QType, filter/FIRFilter.cpp:381)
// and keeps bits [16, 32) of the sum (fromQ, :300).  Everything is therefore arithmetic mod 2^32,
// which byte limbs reproduce exactly:  x = xl + 2^8 xh  (xl unsigned, xh signed),
// tap q = sum_l 2^(8 l) q_l with balanced signed digits q_l in [-128, 127], and
//   x q  =  sum_{dl, l}  2^(8 (dl + l))  x_dl q_l        (terms with 8 (dl + l) >= 32 vanish).
// Each limb product is one int8 Toeplitz GEMM accumulated in int32 (wrapping, like the reference).
struct FirImmaPlan {
    bool ready = false;
    int K = 0;          // taps
    int NB = 0;         // 32-sample k-blocks per output row: ceil((K + 7) / 32)
    int nlt = 2;        // tap limbs (2: |q| < 2^15 - 128, i.e. |h| < 0.498; 3; 4: any int32)
    int dc = 1, tc = 1; // data / tap components (1 real, 2 complex)
    void *d_frag = nullptr;   // [NB][tc][nlt][32] uint2: per-lane B fragments of the tap Toeplitz blocks
    size_t frag_capacity = 0;
};

constexpr size_t kFirImmaMinTaps = 2;      // measured (tools/sweep_i16.sh): ahead of the IMAD tile kernel from 2 taps up
constexpr size_t kFirImmaMaxTaps = 2048;

// Decides whether the tensor-core path serves (dtype, taps, M, L); leaves p.ready false otherwise.
int fir_imma_configure(FirImmaPlan &p, int dtype, const double *taps, size_t ntaps, bool complex_taps, size_t M, size_t L,
                       bool force);
void fir_imma_destroy(FirImmaPlan &p);
int fir_imma_launch(const FirImmaPlan &p, const void *d_in, size_t in_elems, void *d_out, size_t n_out, int sm_count,
                    cudaStream_t stream);

// tcgen05 / TMEM variant of the same algebra (fir_umma.cu): 2-limb taps, built on a ready FirImmaPlan.
struct FirUmmaPlan {
    bool ready = false;
    int K = 0, NB = 0;  // NB: 32-sample k-blocks per 16-output row: ceil((K + 15) / 32)
    int nlt = 2, dc = 1, tc = 1;
    void *d_btile = nullptr;  // [NB][N x 32 B] canonical K-major B tiles, N = 16 tc nlt
    size_t capacity = 0;
};
int fir_umma_configure(FirUmmaPlan &p, const FirImmaPlan &base, const double *taps, bool force);
void fir_umma_destroy(FirUmmaPlan &p);
int fir_umma_launch(const FirUmmaPlan &p, const void *d_in, size_t in_elems, void *d_out, size_t n_out, int sm_count,
                    cudaStream_t stream);

// Second tcgen05 formulation (fir_umma32.cu): 32-byte-swizzled planes, 32 outputs per row, N = 128.
struct FirUmma32Plan {
    bool ready = false;
    int K = 0, NB = 0;  // NB: k-blocks per 32-output row: ceil((K + 31) / 32)
    int dc = 1;
    void *d_bmat = nullptr;   // [dc][NB][N x 32 B] B tiles, N = 32 x dc x 2
    size_t capacity = 0;
    // operand-swapped variant (fir_umma32t_kernel, complex int16 data): tap tiles as the A operand in tensor
    // memory, row-major [dc][NB][128 rows x 32 B]
    bool swapped = false;
    void *d_amat = nullptr;
    size_t a_capacity = 0;
};
int fir_umma32_configure(FirUmma32Plan &p, const FirImmaPlan &base, const double *taps, bool force, int swap = -1);
void fir_umma32_destroy(FirUmma32Plan &p);
int fir_umma32_launch(const FirUmma32Plan &p, const void *d_in, size_t in_elems, void *d_out, size_t n_out, int sm_count,
                      cudaStream_t stream);

// tcgen05 polyphase resampler for int16 streams (fir_ummap.cu): L, M <= 4, 2-digit taps.
struct FirUmmaPPlan {
    bool ready = false;
    int L = 1, M = 1, NB = 0, N = 0, dc = 1, nstage = 1, pstage = 1, npass = 1;
    void *d_bmat = nullptr;   // [M][dc][NB][N x 32 B] B tiles, N = 16 L dc 2
    size_t capacity = 0;
};
int fir_ummap_configure(FirUmmaPPlan &p, int dtype, const double *taps, size_t ntaps, bool complex_taps, size_t M, size_t L, bool force);
void fir_ummap_destroy(FirUmmaPPlan &p);
// nq = output blocks q (= consumed / M); outputs written = nq * L
int fir_ummap_launch(const FirUmmaPPlan &p, const void *d_in, size_t in_elems, void *d_out, size_t nq, int sm_count, cudaStream_t stream);

} // namespace b200c
