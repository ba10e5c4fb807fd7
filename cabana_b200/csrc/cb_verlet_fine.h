// Arguments of the v1 Verlet passes (cb_verlet_fine.cu).
#pragma once

#include <cuda_runtime.h>

#include "cb_common.cuh"

namespace cb
{

struct FineArgs
{
    const float4* q;      // sorted slots: (x-ox, y-oy, z-oz) as float, w = particle id bits;
                          // q[n] is a sentinel at +infinity (list padding)
    const unsigned* ids;  // sorted slot -> particle id (the LCL permutation)
    const double* xs;     // sorted exact coordinates
    const double* ys;
    const double* zs;
    const unsigned* cell_off; // refined cells, [ncell+1]
    Grid ug;                  // the reference's grid (delta = ratio*r): exact band prune
    int lgm;                  // log2 of the refinement factor m per dimension
    int R;                    // reference stencil range in user cells (ceil(1/ratio))
    int nf[3];                // refined cells per dimension
    double hf[3];             // refined cell size
    int K[3];                 // refined cells the cutoff can reach per dimension
    signed char kz[81];       // kz[|da|*9+|db|]: z reach (refined cells) of stencil row
                              // (da,db), -1 if the row is out of reach (host-built,
                              // conservative: uses r*(1+1e-9) + margin)
    double rsqr;              // r*r (Cabana_VerletList.hpp:239)
    double band;              // s > rsqr - band: evaluate the reference's cell prune
    float t_lo, t_hi;         // FP32 filter thresholds r^2 -+ tau
    long long n, begin, end;
    int full_range;           // begin == 0 && end == n
    long long ncell;
    int* counts;
    const int* offsets;
    int* neighbors;
    long long width;
    unsigned* worklist;   // [ncell] cells handed from the column kernel to the general one
    unsigned* work_count; // device counter; nullptr worklist = general kernel only
    // single test pass: rows are appended here in binned order, then reordered
    int* tmp;                   // [tmp_capacity] neighbour ids
    unsigned* tmp_off;          // [n] row start in tmp, per particle id
    int* offsets_direct;        // CB_ROWS_BINNED: offsets[pid] = row start in tmp (no reorder)
    unsigned long long* cursor; // bump allocator of tmp (ids reserved so far)
    int* overflow;              // set when tmp_capacity was too small
    long long tmp_capacity;
};

// Test every pair once: counts[pid], tmp rows, tmp_off[slot].
int launch_fine_single( const FineArgs& a, int algorithm, cudaStream_t stream );
// Move the rows to neighbors[offsets[pid]..] (CSR) or neighbors[pid*width..] (2D).
int launch_fine_reorder( const FineArgs& a, int layout, cudaStream_t stream );

} // namespace cb
