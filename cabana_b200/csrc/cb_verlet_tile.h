// Arguments of the v2 ("tile") Verlet build (cb_verlet_tile.cu).
#pragma once

#include <cuda_runtime.h>

#include "cb_common.cuh"
#include "cb_internal.h"

namespace cb
{

constexpr int kTileHomes = 16;   // home particles per tile (rows of the m16n8k8 tile)
constexpr int kTileCands = 8;    // candidates per MMA tile (columns)
constexpr int kPieceTiles = 16;  // MMA tiles staged per piece (128 candidates)
constexpr int kChunkTiles = 32;  // MMA tiles per mask chunk (one bit per tile in a word)
constexpr int kChunkWords = 128; // 32 lanes x 4 words = 512 B of hit bits per chunk

// Host-side description of the internal ("pencil") grid: columns of about one cutoff in
// x and y, thin cells in z.  It is independent of the user's cell_size_ratio: the
// reference's own grid only enters through the exact band checks (TileArgs::ug, R).
struct TileGrid
{
    cb_grid g;   // internal grid in the library's CartesianGrid form (binning)
    int ncx, ncy, nz;
    int zb;      // z cells per block (tiles never straddle a block)
    int nzb;     // blocks per column
    int kz;      // z cells the cutoff can reach (conservative)
    long long ncols, nblocks, ncells;
};

struct TileArgs
{
    // cell-sorted particles
    const float4* q;          // (x-min, y-min, z-min) as float, w = particle id bits
    const unsigned* permute;  // sorted slot -> particle id
    const unsigned* cell_off; // internal cells, [ncells+1], z fastest
    PosAccess x;              // the user's positions (exact tier reads them in place)
    // internal grid
    int ncx, ncy, nz, zb, nzb, kz;
    float wx, wy, hz;         // cell sizes (tile origins only; any float works)
    // tiles
    const uint4* recs;         // {first slot, col a | b << 16, zlo | zhi << 16, np | #mma tiles << 8}
    const uint2* spans;        // [ntiles][9] candidate spans (start slot, length)
    const int* chunk_off;      // [ntiles+1] first mask chunk of every tile
    const int* ntiles_dev;     // tile_base[nblocks] (device)
    long long mask_capacity;   // chunks the mask buffer holds
    uint4* masks;              // [chunk][lane] hit bits, 512 B per chunk
    unsigned* ticket;          // dynamic tile scheduler
    int* overflow;             // set when mask_capacity is too small
    // decisions
    Grid ug;                   // the reference's grid (delta = ratio * r): exact band prune
    int R;                     // reference stencil range in user cells
    double rsqr, band;
    float r2hi, r2lo;          // tf32 split of r*r
    float tau;                 // |c| <= tau: decided by the exact tier
    // rows
    long long n, begin, end;
    int* counts;
    int* cnt_sorted;       // counts in sorted-slot order (0 for rows that are not built)
    const int* dst_sorted; // CSR: offsets in sorted-slot order (fill pass)
    const int* offsets; // CSR row starts (fill pass); nullptr for 2D
    int* neighbors;
    long long width;    // 2D row width
    // speculative fill (launched before the host has seen the list size): the kernel leaves
    // without writing when the mask buffer overflowed or the list does not fit
    const long long* spec_total; // device: total stored neighbours (nullptr: not speculative)
    long long spec_capacity;     // ints the neighbour array holds
    int* spec_failed;            // set when the kernel left early
    int fill_sort;               // fill pass: expand a lane's mask words largest first
    // diagnostics (cb_verlet_filter_selftest): [0] max |c_mma - c_exact| as float bits,
    // [1] number of values that missed the bound
    unsigned* diag_maxerr;
};

// Internal grid for radius r in box [min,max]; n particles (bounds the cell count).
void make_tile_grid( TileGrid& tg, const double* grid_min, const double* grid_max,
                     double radius, long long n );

// Error bound of the tf32 filter for this grid (DESIGN.md "Exactness"); tau = 2 * bound.
double tile_filter_bound( const TileGrid& tg, double radius, int nzc );

// Fused binning on the internal grid: cell_off[ncells+1] (every column padded to a multiple
// of 8 slots plus 8), q / permute in cell-sorted order (pad slots hold far-away sentinels and
// id -1).  The sorted arrays need sorted_capacity(tg, n) + 8 slots.  n_dev (optional, device):
// the particle count when the host only knows the bound x.n (ghost counts that never left the
// GPU, cb_slab_step).
inline long long sorted_capacity( const TileGrid& tg, long long n ) { return n + 16 * tg.ncols; }
int tile_bin( const TileGrid& tg, const cb_positions& x, int* cell_counts, unsigned* cell_off,
              uint2* cellslot, unsigned char* pads, float4* q, unsigned* permute,
              DeviceBuffer& scan_scratch, cudaStream_t stream,
              const long long* n_dev = nullptr );

// Tile records + mask chunk offsets.  block_tiles/tile_base: [nblocks+1] ints;
// recs: capacity n/16 + nblocks + 1; chunk_off: same + 1.
int tile_plan( const TileGrid& tg, const unsigned* cell_off, bool half, int* block_tiles,
               int* tile_base, uint4* recs, uint2* spans, int* tile_chunks, int* chunk_off,
               long long rec_capacity, DeviceBuffer& scan_scratch, cudaStream_t stream );

int tile_count_pass( const TileArgs& a, bool half, cudaStream_t stream );
int tile_sorted_dst( const TileArgs& a, long long ncells, long long ns_cap, int* dst_sorted,
                     cudaStream_t stream );
int tile_fill_pass( const TileArgs& a, bool csr, cudaStream_t stream );
// Count pass with every filter value checked against the exact arithmetic (tests only).
int tile_diag_pass( const TileArgs& a, bool half, cudaStream_t stream );

} // namespace cb
