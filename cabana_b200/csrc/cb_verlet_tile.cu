// VerletList construction kernels, v2: distance TILES on the tensor cores, bulk-copy
// staging, count pass -> scan -> fill pass (no temporary list, no reorder).
//
// Same reference semantics as cb_verlet.cu (core/src/Cabana_VerletList.hpp:316-474 count,
// :507-562 processCounts, :572-713 fill).  What ncu said about v1 (profiles/r01_*): the
// SIMT distance test (10 instructions per 32 pairs, L1TEX 72 %, issue 60 %) and the
// temporary-then-reorder design (11.6 GB of extra DRAM traffic) cannot reach the roofline
// target.  v2 therefore
//
//  * bins on an internal "pencil" grid (columns of one cutoff in x,y; thin cells in z), so
//    the candidates of 16 consecutive home particles are <= 9 CONTIGUOUS spans of the
//    cell-sorted array;
//  * moves every span into shared memory with cp.async.bulk (the TMA engine, mbarrier
//    completion, double buffered): candidate loads leave the LSU;
//  * evaluates s - r^2 = |xi|^2 + |xj|^2 - 2 xi.xj - r^2 for a tile of 16 home particles x 8
//    candidates with two mma.sync.m16n8k8 tf32 instructions on tile-local coordinates
//    split hi/lo (22 significant bits), the cutoff folded into the product, so the SIGN of
//    the accumulator is the in/out decision: one funnel shift per value records it, one
//    FMNMX tracks min |c|.  Values with |c| <= tau (a rigorous bound on the filter's error,
//    tile_filter_bound) are re-decided by the exact FP64 reference arithmetic, the exact
//    half criterion and the reference's cell prune (Cabana_VerletList.hpp:401-403): every
//    in/out decision is bit-identical to the reference's;
//  * count pass: hit bits -> counts[] + a compact bit mask per tile (~96 B per particle);
//    offsets = exclusive scan of counts in particle order (:478-491); fill pass: expands
//    the masks into rows staged in shared memory and writes every row ONCE, coalesced, at
//    its final place.
#include "cb_verlet_tile.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

namespace cb
{
namespace
{

constexpr int kWarpsT = 8;
constexpr int kBlockT = kWarpsT * 32;
constexpr int kPieceEntries = kPieceTiles * kTileCands; // 128 candidates
constexpr int kChunkEntries = kChunkTiles * kTileCands; // 256 candidates
constexpr int kRowCap = 1280;                           // ids staged per tile in the fill pass

// ---------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk copy (TMA engine) + tf32 mma
// ---------------------------------------------------------------------------------------
CB_D unsigned smem_u32( const void* p )
{
    return (unsigned)__cvta_generic_to_shared( p );
}
CB_D void mbar_init( unsigned long long* b, unsigned count )
{
    asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( smem_u32( b ) ),
                  "r"( count )
                  : "memory" );
}
CB_D void mbar_expect_tx( unsigned long long* b, unsigned bytes )
{
    asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                      smem_u32( b ) ),
                  "r"( bytes )
                  : "memory" );
}
CB_D bool mbar_try_wait( unsigned long long* b, unsigned parity )
{
    unsigned ok;
    asm volatile( "{\n\t.reg .pred p;\n\t"
                  "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                  "selp.u32 %0, 1, 0, p;\n\t}"
                  : "=r"( ok )
                  : "r"( smem_u32( b ) ), "r"( parity )
                  : "memory" );
    return ok != 0u;
}
CB_D void bulk_g2s( void* dst, const void* src, unsigned bytes, unsigned long long* b )
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"( smem_u32( dst ) ),
        "l"( src ), "r"( bytes ), "r"( smem_u32( b ) )
        : "memory" );
}
// Ampere-style per-lane asynchronous copy (LDGSTS): 16 bytes global -> shared, L1-cached.
CB_D void cp_async16( void* dst, const void* src )
{
    asm volatile( "cp.async.ca.shared.global [%0], [%1], 16;" ::"r"( smem_u32( dst ) ), "l"( src )
                  : "memory" );
}
CB_D void cp_async_commit() { asm volatile( "cp.async.commit_group;" ::: "memory" ); }
CB_D void cp_async_wait_all() { asm volatile( "cp.async.wait_group 0;" ::: "memory" ); }
CB_D void fence_proxy_async()
{
    asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" );
}
CB_D void fence_mbar_init()
{
    asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
}

// D = A(16x8, row) * B(8x8, col) + C, tf32 inputs, fp32 accumulate.
// Thread (g = lane>>2, t = lane&3) holds  A: (g,t) (g+8,t) (g,t+4) (g+8,t+4);
// B: (k=t, n=g) (k=t+4, n=g);  C/D: (g,2t) (g,2t+1) (g+8,2t) (g+8,2t+1).
CB_D void mma_tf32( float ( &d )[4], float a0, float a1, float a2, float a3, float b0,
                    float b1, const float ( &c )[4] )
{
    asm( "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 "
         "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
         : "=f"( d[0] ), "=f"( d[1] ), "=f"( d[2] ), "=f"( d[3] )
         : "r"( __float_as_uint( a0 ) ), "r"( __float_as_uint( a1 ) ),
           "r"( __float_as_uint( a2 ) ), "r"( __float_as_uint( a3 ) ),
           "r"( __float_as_uint( b0 ) ), "r"( __float_as_uint( b1 ) ), "f"( c[0] ),
           "f"( c[1] ), "f"( c[2] ), "f"( c[3] ) );
}

// Keep the 11 significant bits tf32 has: exactly representable, so the tensor core's own
// fp32 -> tf32 conversion (whatever its rounding) leaves the value alone.
CB_D float trunc_tf32( float v )
{
    return __uint_as_float( __float_as_uint( v ) & 0xffffe000u );
}

struct TileRec
{
    unsigned first;
    int ca, cb, np, zlo, zhi, ntile8;
};
// {first slot, column a | b << 16, zlo | zhi << 16, np | #mma tiles << 8}
CB_D TileRec unpack_rec( const uint4 r )
{
    TileRec t;
    t.first = r.x;
    t.ca = (int)( r.y & 0xffffu );
    t.cb = (int)( r.y >> 16 );
    t.zlo = (int)( r.z & 0xffffu );
    t.zhi = (int)( r.z >> 16 );
    t.np = (int)( r.w & 0xffu );
    t.ntile8 = (int)( r.w >> 8 );
    return t;
}

// Span s (0..8) of the candidates of a tile in column (ca,cb), z cells [zlo,zhi]:
// s = 0..2: same x column (db = -1,0,1); 3..5: da = +1; 6..8: da = -1.  Half lists skip
// da = -1: the cell index is monotone in x, so every particle there has a smaller x and
// fails isValid (Cabana_NeighborList.hpp:139-149).
struct GridInts
{
    int ncx, ncy, nz, kz;
};
CB_D void span_of( const GridInts& gi, const unsigned* __restrict__ cell_off, int ca,
                   int cb_, int zlo, int zhi, bool half, int s, unsigned& start,
                   unsigned& len )
{
    start = 0u;
    len = 0u;
    const int grp = s / 3;
    const int da = grp == 0 ? 0 : ( grp == 1 ? 1 : -1 );
    const int db = s - grp * 3 - 1;
    const int aa = ca + da, bb = cb_ + db;
    if ( aa < 0 || aa >= gi.ncx || bb < 0 || bb >= gi.ncy || ( half && da < 0 ) )
        return;
    const int z0 = max( zlo - gi.kz, 0 );
    const int z1 = min( zhi + gi.kz, gi.nz - 1 );
    const long long c0 = ( (long long)aa * gi.ncy + bb ) * gi.nz;
    // the window is widened to multiples of 8 slots: what it gains are particles of cells
    // the cutoff cannot reach, or the column's sentinel pad -- never a neighbour
    start = cell_off[c0 + z0] & ~7u;
    len = ( ( cell_off[c0 + z1 + 1] + 7u ) & ~7u ) - start;
}

// The reference's cell-level prune for the pair (p, n), evaluated exactly on the USER grid
// (Cabana_VerletList.hpp:401-403 -> impl/Cabana_CartesianGrid.hpp:207-223), plus the
// stencil range check (Cabana_LinkedCellList.hpp:112-118).
__device__ __noinline__ bool reference_prune_passes( const Grid& ug, int R, double rsqr,
                                                     double xp, double yp, double zp,
                                                     double xn, double yn, double zn )
{
    int cn[3];
    const double pp[3] = { xp, yp, zp };
    const double pn[3] = { xn, yn, zn };
#pragma unroll
    for ( int d = 0; d < 3; ++d )
    {
        const int cp = min( max( locate_1d( ug, d, pp[d] ), 0 ), ug.nx[d] - 1 );
        cn[d] = min( max( locate_1d( ug, d, pn[d] ), 0 ), ug.nx[d] - 1 );
        if ( abs( cn[d] - cp ) > R )
            return false;
    }
    return min_distance_sq( ug, xp, yp, zp, cn[0], cn[1], cn[2] ) <= rsqr;
}

// Exact tier: the reference's own arithmetic for the pair (particle pi, candidate pj).
template <bool HALF>
__device__ __noinline__ bool exact_decide( const TileArgs& a, int pi, int pj )
{
    if ( pi < 0 || pj < 0 || pi == pj )
        return false;
    const long long oi = a.x.offset( pi ), oj = a.x.offset( pj );
    const double xp = a.x.base[oi], yp = a.x.base[oi + a.x.comp_stride],
                 zp = a.x.base[oi + 2 * a.x.comp_stride];
    const double xn = a.x.base[oj], yn = a.x.base[oj + a.x.comp_stride],
                 zn = a.x.base[oj + 2 * a.x.comp_stride];
    bool ok = true;
    if ( HALF )
        ok = half_criterion( xp, yp, zp, xn, yn, zn );
    const double d2 = pair_dist_sq( xp, yp, zp, xn, yn, zn );
    ok = ok && ( d2 <= a.rsqr ); // closed cutoff (:254)
    if ( ok && d2 > a.rsqr - a.band )
        ok = reference_prune_passes( a.ug, a.R, a.rsqr, xp, yp, zp, xn, yn, zn );
    return ok;
}

__device__ __noinline__ double exact_c( const TileArgs& a, int pi, int pj )
{
    const long long oi = a.x.offset( pi ), oj = a.x.offset( pj );
    const double d2 = pair_dist_sq(
        a.x.base[oi], a.x.base[oi + a.x.comp_stride], a.x.base[oi + 2 * a.x.comp_stride],
        a.x.base[oj], a.x.base[oj + a.x.comp_stride], a.x.base[oj + 2 * a.x.comp_stride] );
    return d2 - a.rsqr;
}

// ---------------------------------------------------------------------------------------
// binning on the internal grid (LinkedCellList::build semantics, Cabana_LinkedCellList.hpp
// :651-739, fused for this consumer): positions are read from the user's slice TWICE in
// coalesced order -- (1) locate + warp-aggregated atomic that both histograms and claims the
// slot, (2) after the scan, scatter the FP32 origin-relative record straight to its sorted
// slot -- instead of histogram, permute fill and a random gather.
// Every column is padded with sentinel slots to a multiple of 8 plus 8, so 8-aligned windows
// of the sorted array never leave their column.
// ---------------------------------------------------------------------------------------
constexpr int kBinBatch = 4; // particles per thread and round: their loads are issued together

__global__ void __launch_bounds__( 256 )
    k_tbin_count( PosAccess x, Grid g, long long n, int* __restrict__ counts,
                  uint2* __restrict__ cellslot, const long long* __restrict__ n_dev )
{
    if ( n_dev )
        n = min( n, *n_dev );
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    for ( long long p0 = (long long)blockIdx.x * ( 256 * kBinBatch ); p0 < n;
          p0 += (long long)gridDim.x * ( 256 * kBinBatch ) )
    {
        double px[kBinBatch], py[kBinBatch], pz[kBinBatch];
#pragma unroll
        for ( int k = 0; k < kBinBatch; ++k )
        {
            const long long p = p0 + k * 256 + threadIdx.x;
            px[k] = py[k] = pz[k] = 0.0;
            if ( p < n )
            {
                const long long off = x.offset( p );
                px[k] = x.base[off];
                py[k] = x.base[off + x.comp_stride];
                pz[k] = x.base[off + 2 * x.comp_stride];
            }
        }
#pragma unroll
        for ( int k = 0; k < kBinBatch; ++k )
        {
            const long long p = p0 + k * 256 + threadIdx.x;
            const bool valid = p < n;
            int c = -1 - (int)lane; // unique dummy key for idle lanes
            if ( valid )
            {
                int ci = locate_1d( g, 0, px[k] );
                int cj = locate_1d( g, 1, py[k] );
                int ck = locate_1d( g, 2, pz[k] );
                // points outside [min,max] are undefined behaviour in the reference; clamp
                ci = min( max( ci, 0 ), g.nx[0] - 1 );
                cj = min( max( cj, 0 ), g.nx[1] - 1 );
                ck = min( max( ck, 0 ), g.nx[2] - 1 );
                c = cardinal_index( g, ci, cj, ck );
            }
            // Warp aggregation only over RUNS of equal cells (neighbouring lanes): one shuffle and
            // one ballot instead of __match_any_sync, whose all-pairs comparison stalled this
            // kernel on the MIO queue (ncu: short-scoreboard 25 cycles per issue).  Particle arrays
            // that are sorted or lattice-ordered put equal cells next to each other; unordered
            // input degenerates to one atomic per particle, which is what it needs anyway.
            const int cprev = __shfl_up_sync( kFullMask, c, 1 );
            const bool head = lane == 0u || cprev != c;
            const unsigned heads = __ballot_sync( kFullMask, head );
            const unsigned below = heads & ( lt | ( 1u << lane ) );     // heads at or below me
            const int leader = 31 - __clz( below );                    // start of my run
            const unsigned above = heads & ~( lt | ( 1u << lane ) );    // heads above me
            const int run_end = above ? __ffs( above ) - 1 : 32;        // one past my run
            int base = 0;
            if ( valid && head )
                base = atomicAdd( &counts[c], run_end - (int)lane );
            base = __shfl_sync( kFullMask, base, leader );
            if ( valid )
                cellslot[p] = make_uint2( (unsigned)c, (unsigned)( base + (int)lane - leader ) );
        }
    }
}

// One warp per column: pad = slots added behind the column's last cell.
__global__ void __launch_bounds__( 256 )
    k_tbin_pad( int* __restrict__ counts, long long ncols, int nz,
                unsigned char* __restrict__ pads )
{
    const unsigned lane = lane_id();
    const long long warp = ( (long long)blockIdx.x * 256 + threadIdx.x ) >> 5;
    const long long nwarps = ( (long long)gridDim.x * 256 ) >> 5;
    for ( long long col = warp; col < ncols; col += nwarps )
    {
        int* cc = counts + col * nz;
        int sum = 0;
        for ( int k = (int)lane; k < nz; k += 32 )
            sum += cc[k];
        sum = warp_reduce_sum( sum );
        if ( lane == 0 )
        {
            const int pad = ( ( 8 - ( sum & 7 ) ) & 7 ) + 8;
            pads[col] = (unsigned char)pad;
            cc[nz - 1] += pad;
        }
    }
}

__global__ void __launch_bounds__( 256 )
    k_tbin_scatter( PosAccess x, long long n, const uint2* __restrict__ cellslot,
                    const unsigned* __restrict__ cell_off, float4* __restrict__ q,
                    unsigned* __restrict__ permute, double ox, double oy, double oz,
                    const long long* __restrict__ n_dev )
{
    if ( n_dev )
        n = min( n, *n_dev );
    for ( long long p0 = (long long)blockIdx.x * ( 256 * kBinBatch ); p0 < n;
          p0 += (long long)gridDim.x * ( 256 * kBinBatch ) )
    {
        double px[kBinBatch], py[kBinBatch], pz[kBinBatch];
        uint2 cs[kBinBatch];
#pragma unroll
        for ( int k = 0; k < kBinBatch; ++k )
        {
            const long long p = p0 + k * 256 + threadIdx.x;
            px[k] = py[k] = pz[k] = 0.0;
            cs[k] = make_uint2( 0u, 0u );
            if ( p < n )
            {
                const long long off = x.offset( p );
                px[k] = x.base[off];
                py[k] = x.base[off + x.comp_stride];
                pz[k] = x.base[off + 2 * x.comp_stride];
                cs[k] = cellslot[p];
            }
        }
        unsigned base[kBinBatch];
#pragma unroll
        for ( int k = 0; k < kBinBatch; ++k )
            base[k] = ( p0 + k * 256 + threadIdx.x < n ) ? cell_off[cs[k].x] : 0u;
#pragma unroll
        for ( int k = 0; k < kBinBatch; ++k )
        {
            const long long p = p0 + k * 256 + threadIdx.x;
            if ( p < n )
            {
                const unsigned s = base[k] + cs[k].y;
                q[s] = make_float4( __double2float_rn( px[k] - ox ), __double2float_rn( py[k] - oy ),
                                    __double2float_rn( pz[k] - oz ), __int_as_float( (int)p ) );
                permute[s] = (unsigned)p;
            }
        }
    }
}

// Sentinels: the pad slots of every column and 8 slots behind the end of the array.
__global__ void __launch_bounds__( 256 )
    k_tbin_sentinels( const unsigned* __restrict__ cell_off, const unsigned char* __restrict__ pads,
                      long long ncols, int nz, float4* __restrict__ q,
                      unsigned* __restrict__ permute )
{
    const float4 sent = make_float4( 1.0e18f, 1.0e18f, 1.0e18f, __int_as_float( -1 ) );
    for ( long long col = (long long)blockIdx.x * 256 + threadIdx.x; col <= ncols;
          col += (long long)gridDim.x * 256 )
    {
        const unsigned end = cell_off[min( col + 1, ncols ) * nz];
        const int pad = col < ncols ? (int)pads[col] : 0;
        if ( col == ncols )
        {
            for ( int k = 0; k < 8; ++k )
            {
                q[end + k] = sent;
                permute[end + k] = 0xffffffffu;
            }
            continue;
        }
        for ( int k = 1; k <= pad; ++k )
        {
            q[end - k] = sent;
            permute[end - k] = 0xffffffffu;
        }
    }
}

// ---------------------------------------------------------------------------------------
// plan: tiles of <= 16 consecutive particles inside one block (column x zb cells)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__( 256 )
    k_plan_blocks( const unsigned* __restrict__ cell_off, int nz, int zb, int nzb,
                   long long nblocks, int* __restrict__ block_tiles )
{
    for ( long long b = (long long)blockIdx.x * 256 + threadIdx.x; b < nblocks;
          b += (long long)gridDim.x * 256 )
    {
        const long long col = b / nzb;
        const int z0 = (int)( b - col * nzb ) * zb;
        const int z1 = min( z0 + zb, nz );
        const long long base = col * nz;
        const unsigned nb = cell_off[base + z1] - cell_off[base + z0];
        block_tiles[b] = (int)( ( nb + kTileHomes - 1 ) / kTileHomes );
    }
}

// Eight threads per block share its tiles (thread `sub` takes tiles sub, sub + 8, ...).
constexpr int kPlanSplit = 8;
__global__ void __launch_bounds__( 128 )
    k_plan_tiles( const unsigned* __restrict__ cell_off, GridInts gi, int zb, int nzb,
                  long long nblocks, int half, const int* __restrict__ block_tiles,
                  const int* __restrict__ tile_base, uint4* __restrict__ recs,
                  uint2* __restrict__ spans, int* __restrict__ tile_chunks,
                  long long rec_capacity )
{
    for ( long long item = (long long)blockIdx.x * 128 + threadIdx.x; item < nblocks * kPlanSplit;
          item += (long long)gridDim.x * 128 )
    {
        const long long b = item / kPlanSplit;
        const int sub = (int)( item - b * kPlanSplit );
        const int nt = block_tiles[b];
        if ( sub >= nt )
            continue;
        const long long tb = tile_base[b];
        const long long col = b / nzb;
        const int z0 = (int)( b - col * nzb ) * zb;
        const int z1 = min( z0 + zb, gi.nz );
        const long long base = col * gi.nz;
        const int ca = (int)( col / gi.ncy );
        const int cb_ = (int)( col - (long long)ca * gi.ncy );
        const unsigned p0 = cell_off[base + z0];
        const unsigned pend = cell_off[base + z1];
        int zc = z0;
        for ( int ti = sub; ti < nt; ti += kPlanSplit )
        {
            const unsigned first = p0 + (unsigned)( ti * kTileHomes );
            const unsigned last = min( first + (unsigned)kTileHomes, pend ) - 1u;
            while ( cell_off[base + zc + 1] <= first )
                ++zc;
            const int zlo = zc;
            int zz = zc;
            while ( cell_off[base + zz + 1] <= last )
                ++zz;
            const int zhi = zz;
            int T = 0;
            const bool fits = tb + ti < rec_capacity;
            for ( int s = 0; s < 9; ++s )
            {
                unsigned st, len;
                span_of( gi, cell_off, ca, cb_, zlo, zhi, half != 0, s, st, len );
                T += (int)( ( len + kTileCands - 1 ) / kTileCands );
                if ( fits )
                    spans[( tb + ti ) * 9 + s] = make_uint2( st, len );
            }
            if ( fits )
            {
                recs[tb + ti] = make_uint4(
                    first, (unsigned)ca | ( (unsigned)cb_ << 16 ),
                    (unsigned)zlo | ( (unsigned)zhi << 16 ),
                    ( last - first + 1u ) | ( (unsigned)T << 8 ) );
                tile_chunks[tb + ti] = ( T + kChunkTiles - 1 ) / kChunkTiles;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// count pass
// ---------------------------------------------------------------------------------------
constexpr int kTableTiles = 256; // mma tiles whose source slot is tabulated at a time

// Candidate features of one piece, in two planes so that both the stores (one float4 per lane
// and plane: 512 contiguous bytes per instruction) and the B-operand loads are free of bank
// conflicts:  plane A: entry e -> (x hi, x lo, y hi, y lo);  plane B: (z hi, z lo, |x|^2 hi, lo).
// Thread (g,t) of mma tile j reads ONE 8-byte word: feature t of entry 8j + g, i.e.
// plane (t >> 1), float2 index 2 (8j + g) + (t & 1).  Plane B starts 64 bytes past a multiple of
// 128, so the two planes of a half-warp's words fall into different banks.
constexpr int kFeatPlaneB = 2 * kPieceEntries + 8;      // float2 index of plane B
constexpr int kFeatWords = 4 * kPieceEntries + 8;       // float2 words of one feature buffer
CB_D void store_features( float2* feat, int e, float x, float y, float z )
{
    const float hx = trunc_tf32( x ), hy = trunc_tf32( y ), hz = trunc_tf32( z );
    const float n = fmaf( z, z, fmaf( y, y, x * x ) );
    const float nh = trunc_tf32( n );
    *reinterpret_cast<float4*>( feat + 2 * e ) = make_float4( hx, x - hx, hy, y - hy );
    *reinterpret_cast<float4*>( feat + kFeatPlaneB + 2 * e ) = make_float4( hz, z - hz, nh, n - nh );
}

// Everything a warp keeps about its 16 home particles while it sweeps their candidates.
struct HomeTile
{
    float a1[4], a2[4];               // A operands: (g,t) (g+8,t) (g,t+4) (g+8,t+4)
    float cinit[4];                   // accumulator init: |x_i|^2 - r^2 of rows g, g+8
    float hx_g, hx_g8;                // tile-local x of rows g, g+8 (half criterion)
    int pid_g, pid_g8;
    bool act_g, act_g8;
    unsigned selfbase;                // list entry of home particle 0
    unsigned m0, m1, m2, m3;          // hit bits of the current chunk
    float ma;                         // min |c| since the last exact-tier check
    int cnt_g, cnt_g8;
};

// Home operands from the tile-local coordinates of lanes 0..15 (scratch: 288 floats).
//   A1 = [-2H, 1 | -2H, 1]   A2 = [-2L, 0 | -2L, 0]   C = |x_i|^2 - r^2
CB_D void load_home_operands( HomeTile& H, float* scr, float xh, float yh, float zh, int pid,
                              unsigned actmask, float r2hi, float r2lo, unsigned lane )
{
    const int g = (int)( lane >> 2 ), t = (int)( lane & 3u );
    __syncwarp();
    if ( lane < 16u )
    {
        const float Hx = trunc_tf32( xh ), Hy = trunc_tf32( yh ), Hz = trunc_tf32( zh );
        const float N = fmaf( zh, zh, fmaf( yh, yh, xh * xh ) );
        // (the two halves of A1 / A2 and the accumulator pairs hold equal values; they are
        // kept in separate words so every operand register is loaded on its own instead of
        // being re-copied inside the sweep loop)
        float4* f = reinterpret_cast<float4*>( scr + 16 * lane );
        f[0] = make_float4( -2.f * Hx, -2.f * Hy, -2.f * Hz, 1.f );
        f[1] = f[0];
        f[2] = make_float4( -2.f * ( xh - Hx ), -2.f * ( yh - Hy ), -2.f * ( zh - Hz ), 0.f );
        f[3] = f[2];
        // (N - r2hi) - r2lo: N and r2hi are close in magnitude, r2lo is the small part
        scr[256 + 2 * lane] = scr[256 + 2 * lane + 1] = ( N - r2hi ) - r2lo;
    }
    __syncwarp();
    H.a1[0] = scr[16 * g + t];
    H.a1[1] = scr[16 * ( g + 8 ) + t];
    H.a1[2] = scr[16 * g + 4 + t];
    H.a1[3] = scr[16 * ( g + 8 ) + 4 + t];
    H.a2[0] = scr[16 * g + 8 + t];
    H.a2[1] = scr[16 * ( g + 8 ) + 8 + t];
    H.a2[2] = scr[16 * g + 12 + t];
    H.a2[3] = scr[16 * ( g + 8 ) + 12 + t];
    H.cinit[0] = scr[256 + 2 * g];
    H.cinit[1] = scr[256 + 2 * g + 1];
    H.cinit[2] = scr[256 + 2 * ( g + 8 )];
    H.cinit[3] = scr[256 + 2 * ( g + 8 ) + 1];
    H.hx_g = __shfl_sync( kFullMask, xh, g );
    H.hx_g8 = __shfl_sync( kFullMask, xh, g + 8 );
    H.pid_g = __shfl_sync( kFullMask, pid, g );
    H.pid_g8 = __shfl_sync( kFullMask, pid, g + 8 );
    H.act_g = ( actmask >> g ) & 1u;
    H.act_g8 = ( actmask >> ( g + 8 ) ) & 1u;
    H.m0 = H.m1 = H.m2 = H.m3 = 0u;
    H.ma = 3.0e38f;
    H.cnt_g = H.cnt_g8 = 0;
    __syncwarp();
}

// `seg` (<= 32) mma tiles whose features start at feat2 (tile j at feat2 + 32 j): the sign of
// s - r^2 goes into the hit words (tile j ends up at bit seg-1-j), min |c| into H.ma; then
// the exact tier for values inside the error band, and the x-major half criterion for the
// first n_filter tiles (those of the home x column).  cpid: particle ids of the same entries.
template <bool HALF, bool DIAG, int CSTRIDE>
CB_D void sweep_segment( const TileArgs& a, HomeTile& H, const float2* feat2, const int* cpid,
                         int seg, int n_filter, float tau, unsigned lane )
{
    const int t = (int)( lane & 3u );
    const float2* fp = feat2 + ( t >> 1 ) * kFeatPlaneB + 2 * (int)( lane >> 2 ) + ( t & 1 );
#pragma unroll 4
    for ( int j = 0; j < seg; ++j )
    {
        const float2 b = fp[16 * j];
        float c1[4], c[4];
        mma_tf32( c1, H.a1[0], H.a1[1], H.a1[2], H.a1[3], b.x, b.y, H.cinit );
        mma_tf32( c, H.a2[0], H.a2[1], H.a2[2], H.a2[3], b.x, b.y, c1 );
        H.m0 = __funnelshift_l( __float_as_uint( c[0] ), H.m0, 1 );
        H.m1 = __funnelshift_l( __float_as_uint( c[1] ), H.m1, 1 );
        H.m2 = __funnelshift_l( __float_as_uint( c[2] ), H.m2, 1 );
        H.m3 = __funnelshift_l( __float_as_uint( c[3] ), H.m3, 1 );
        H.ma = fminf( fminf( H.ma, fabsf( c[0] ) ), fabsf( c[1] ) );
        H.ma = fminf( fminf( H.ma, fabsf( c[2] ) ), fabsf( c[3] ) );
    }
    // ---- exact tier: values inside the filter's error band ----------------------------
    if ( DIAG || __any_sync( kFullMask, H.ma <= tau ) )
    {
        for ( int j = 0; j < seg; ++j )
        {
            const float2 b = fp[16 * j];
            float c1[4], c[4];
            mma_tf32( c1, H.a1[0], H.a1[1], H.a1[2], H.a1[3], b.x, b.y, H.cinit );
            mma_tf32( c, H.a2[0], H.a2[1], H.a2[2], H.a2[3], b.x, b.y, c1 );
            const unsigned bit = 1u << ( seg - 1 - j );
#pragma unroll
            for ( int k = 0; k < 4; ++k )
            {
                const int e = kTileCands * j + 2 * t + ( k & 1 );
                const int hp = k < 2 ? H.pid_g : H.pid_g8;
                if ( DIAG )
                {
                    const int cp = cpid[CSTRIDE * e];
                    if ( hp >= 0 && cp >= 0 )
                    {
                        // The bound is proven for s <= 4 r^2 (c <= 3 r^2); farther pairs only
                        // need the right sign, and every value outside +-tau must have it.
                        const double ce = exact_c( a, hp, cp );
                        const float err = (float)fabs( (double)c[k] - ce );
                        if ( ce <= 3.0 * a.rsqr )
                        {
                            atomicMax( a.diag_maxerr, __float_as_uint( err ) );
                            if ( err > tau * 0.5f )
                                atomicAdd( a.diag_maxerr + 1, 1u );
                        }
                        if ( fabsf( c[k] ) > tau && ( c[k] < 0.f ) != ( ce <= 0.0 ) )
                            atomicAdd( a.diag_maxerr + 2, 1u );
                    }
                }
                if ( fabsf( c[k] ) <= tau )
                {
                    const bool hit = exact_decide<HALF>( a, hp, cpid[CSTRIDE * e] );
                    unsigned& m = k == 0 ? H.m0 : ( k == 1 ? H.m1 : ( k == 2 ? H.m2 : H.m3 ) );
                    m = hit ? ( m | bit ) : ( m & ~bit );
                }
            }
        }
    }
    H.ma = 3.0e38f;
    // ---- half lists: x-major criterion on the windows of the home x column -----------
    if ( HALF && n_filter > 0 )
    {
        const unsigned low = seg >= 32 ? 0xffffffffu : ( ( 1u << seg ) - 1u );
#pragma unroll
        for ( int k = 0; k < 4; ++k )
        {
            unsigned& m = k == 0 ? H.m0 : ( k == 1 ? H.m1 : ( k == 2 ? H.m2 : H.m3 ) );
            const float hx = k < 2 ? H.hx_g : H.hx_g8;
            const int hp = k < 2 ? H.pid_g : H.pid_g8;
            unsigned mm = m & low;
            while ( mm )
            {
                const int b = 31 - __clz( mm );
                mm &= ~( 1u << b );
                if ( seg - 1 - b >= n_filter )
                    continue; // da = +1: a larger cell index means a larger x
                const int e = kTileCands * ( seg - 1 - b ) + 2 * t + ( k & 1 );
                const float2 fx = feat2[2 * e];
                const float cx = fx.x + fx.y; // == x - Ox exactly
                bool keep;
                if ( cx > hx )
                    keep = true; // float rounding is monotone
                else if ( cx < hx )
                    keep = false;
                else
                    keep = exact_decide<HALF>( a, hp, cpid[CSTRIDE * e] );
                if ( !keep )
                    m &= ~( 1u << b );
            }
        }
    }
}

// End of a chunk of `in_chunk` (<= 32) tiles: normalise the words (tile i at bit 31 - i),
// drop j == i, count, and store the 512-byte mask record.
template <bool DIAG>
CB_D void flush_chunk( const TileArgs& a, HomeTile& H, int in_chunk, int chunk_i, size_t chunk0,
                       unsigned lane )
{
    const int g = (int)( lane >> 2 ), t = (int)( lane & 3u );
    if ( in_chunk < kChunkTiles )
    {
        const int sh = kChunkTiles - in_chunk;
        H.m0 <<= sh;
        H.m1 <<= sh;
        H.m2 <<= sh;
        H.m3 <<= sh;
    }
    // j != i: the home particles sit at list entries selfbase .. selfbase + 15
    if ( (int)( H.selfbase / kChunkEntries ) <= chunk_i &&
         (int)( ( H.selfbase + 15u ) / kChunkEntries ) >= chunk_i )
    {
#pragma unroll
        for ( int hh = 0; hh < 2; ++hh )
        {
            const unsigned p = H.selfbase + (unsigned)( g + 8 * hh );
            const unsigned e = p % kChunkEntries;
            if ( (int)( p / kChunkEntries ) == chunk_i && (int)( ( e & 7u ) >> 1 ) == t )
            {
                const unsigned keep = ~( 0x80000000u >> ( e >> 3 ) );
                if ( hh == 0 )
                {
                    if ( e & 1u )
                        H.m1 &= keep;
                    else
                        H.m0 &= keep;
                }
                else
                {
                    if ( e & 1u )
                        H.m3 &= keep;
                    else
                        H.m2 &= keep;
                }
            }
        }
    }
    if ( !H.act_g )
        H.m0 = H.m1 = 0u;
    if ( !H.act_g8 )
        H.m2 = H.m3 = 0u;
    H.cnt_g += __popc( H.m0 ) + __popc( H.m1 );
    H.cnt_g8 += __popc( H.m2 ) + __popc( H.m3 );
    if ( !DIAG )
        a.masks[( chunk0 + (size_t)chunk_i ) * 32u + lane] = make_uint4( H.m0, H.m1, H.m2, H.m3 );
    H.m0 = H.m1 = H.m2 = H.m3 = 0u;
}

// ---- per-warp kernel: every warp stages the candidates of its own tile ------------------
// Shared memory of one warp.  TMA staging: candidates land in `raw` by cp.async.bulk.
// LDG staging: every lane fetches its candidates itself through a per-tile source table.
struct __align__( 16 ) CountSmemTma
{
    float2 feat[kFeatWords];      // (also the scratch of load_home_operands between tiles)
    float4 raw[2][kPieceEntries]; // bulk-copy landing zone (q records), double buffered; the
                                  // records' w words are the particle ids the exact tier reads
    unsigned sp_start[16], sp_len[16], sp_pos[16];
    unsigned long long mbar[2];
};
struct __align__( 16 ) CountSmemLdg
{
    float2 feat[kFeatWords];     // (also the scratch of load_home_operands between tiles)
    float4 raw[kPieceEntries];   // cp.async landing zone of the NEXT piece
    int cpid[kPieceEntries];     // particle ids of the staged candidates (exact tier)
    unsigned tsrc[kTableTiles];  // sorted slot of the first entry of every mma tile
};

// Stage piece `pc` (candidate list entries [128 pc, 128 pc + 128)) into raw[buf]: every
// window that intersects the piece is one bulk copy (16-byte records, 128-byte multiples).
// All lanes call; the buffer's previous readers finished before the preceding __syncwarp.
CB_D void issue_piece( CountSmemTma& S, const float4* __restrict__ q, int pc, int buf,
                       unsigned lane )
{
    const unsigned p0 = (unsigned)( pc * kPieceEntries );
    const unsigned p1 = p0 + kPieceEntries;
    unsigned lo = 0u, hi = 0u, p = 0u;
    if ( lane < 9u )
    {
        p = S.sp_pos[lane];
        lo = max( p, p0 );
        hi = min( p + S.sp_len[lane], p1 );
        if ( hi < lo )
            hi = lo;
    }
    const unsigned bytes = ( hi - lo ) * 16u;
    const unsigned total = __reduce_add_sync( kFullMask, bytes );
    if ( lane == 0u )
        mbar_expect_tx( &S.mbar[buf], total );
    __syncwarp();
    if ( bytes )
        bulk_g2s( &S.raw[buf][lo - p0], q + S.sp_start[lane] + ( lo - p ), bytes,
                  &S.mbar[buf] );
}

// Source table of the mma tiles [seg0, seg0 + kTableTiles): the lane that owns window s
// writes the entries of its tiles (tile k of the window starts at sorted slot start + 8k).
CB_D void build_tile_table( unsigned* tsrc, int seg0, int tiles_before, int my_nt,
                            unsigned sp_start )
{
    const int i0 = max( 0, seg0 - tiles_before );
    const int i1 = min( my_nt, seg0 + kTableTiles - tiles_before );
    for ( int i = i0; i < i1; ++i )
        tsrc[tiles_before + i - seg0] = sp_start + (unsigned)( kTileCands * i );
}

template <bool HALF, bool DIAG, bool TMA>
__global__ void __launch_bounds__( kBlockT, 3 )
    k_tile_count( const __grid_constant__ TileArgs a )
{
    using SM = typename std::conditional<TMA, CountSmemTma, CountSmemLdg>::type;
    extern __shared__ __align__( 16 ) unsigned char s_dyn[];
    const unsigned lane = threadIdx.x & 31u;
    const int wib = threadIdx.x >> 5;
    SM& S = reinterpret_cast<SM*>( s_dyn )[wib];
    unsigned phase = 0u; // bit b: parity the next wait on mbar[b] uses
    if constexpr ( TMA )
    {
        if ( lane == 0u )
        {
            mbar_init( &S.mbar[0], 1u );
            mbar_init( &S.mbar[1], 1u );
        }
        fence_mbar_init();
        fence_proxy_async();
        __syncwarp();
    }
    const int ntiles = *a.ntiles_dev;

    // Tiles are handed out by a global ticket, two tickets ahead: the record of the next tile
    // is in flight while the current one is swept.
    auto grab = [&]() -> unsigned
    {
        unsigned v = 0u;
        if ( lane == 0u )
            v = atomicAdd( a.ticket, 1u );
        return __shfl_sync( kFullMask, v, 0 );
    };
    unsigned tile = grab();
    unsigned tile_next = grab();
    uint4 rec_raw = make_uint4( 0u, 0u, 0u, 0u ), rec_next = rec_raw;
    int chunk0 = 0, chunk0_next = 0;
    if ( tile < (unsigned)ntiles )
    {
        rec_raw = a.recs[tile];
        chunk0 = a.chunk_off[tile];
    }
    unsigned tile_nn = 0u;
    for ( ;; tile = tile_next, tile_next = tile_nn, rec_raw = rec_next, chunk0 = chunk0_next )
    {
        if ( tile >= (unsigned)ntiles )
            break;
        tile_nn = grab(); // needed only at the end of this iteration
        if ( tile_next < (unsigned)ntiles )
        {
            rec_next = a.recs[tile_next];
            chunk0_next = a.chunk_off[tile_next];
        }
        const TileRec rc = unpack_rec( rec_raw );

        // ---- home particles (lanes 0..15) ------------------------------------------
        float4 hq = make_float4( -1.0e18f, -1.0e18f, -1.0e18f, __int_as_float( -1 ) );
        if ( (int)lane < rc.np )
            hq = a.q[rc.first + lane];
        const int pid = __float_as_int( hq.w );
        const bool act = (int)lane < rc.np && pid >= a.begin && pid < a.end;
        const unsigned actmask = __ballot_sync( kFullMask, act );
        if ( actmask == 0u )
            continue; // no row of this tile is built (:340)
        if ( (long long)chunk0 + ( rc.ntile8 + kChunkTiles - 1 ) / kChunkTiles >
             a.mask_capacity )
        {
            if ( lane == 0u )
                *a.overflow = 1; // the host grows the mask buffer and reruns
            continue;
        }
        const float Ox = ( (float)rc.ca + 0.5f ) * a.wx;
        const float Oy = ( (float)rc.cb + 0.5f ) * a.wy;
        const float Oz = ( 0.5f * (float)( rc.zlo + rc.zhi ) + 0.5f ) * a.hz;
        // candidate windows (planned): lane s < 9 owns window s
        uint2 sp = make_uint2( 0u, 0u );
        if ( lane < 9u )
            sp = a.spans[(size_t)tile * 9u + lane];
        HomeTile H;
        load_home_operands( H, reinterpret_cast<float*>( S.feat ), hq.x - Ox, hq.y - Oy, hq.z - Oz, pid, actmask, a.r2hi,
                            a.r2lo, lane );
        const int my_nt = (int)( sp.y / kTileCands );
        int incl = my_nt;
#pragma unroll
        for ( int o = 1; o < 16; o <<= 1 )
        {
            const int y = __shfl_up_sync( kFullMask, incl, o );
            if ( (int)lane >= o )
                incl += y;
        }
        const int T = __shfl_sync( kFullMask, incl, 15 );
        const int T0 = __shfl_sync( kFullMask, incl, 2 ); // tiles of the da = 0 windows
        const int tiles_before = incl - my_nt;
        // list entry of home particle h: selfbase + h (the home window is s = 1)
        H.selfbase =
            __shfl_sync( kFullMask, (unsigned)tiles_before * kTileCands - sp.x, 1 ) + rc.first;
        if constexpr ( TMA )
        {
            if ( lane < 16u )
            {
                S.sp_start[lane] = sp.x;
                S.sp_len[lane] = sp.y;
                S.sp_pos[lane] = (unsigned)tiles_before * kTileCands;
            }
            __syncwarp();
            issue_piece( S, a.q, 0, 0, lane );
        }
        else
        {
            build_tile_table( S.tsrc, 0, tiles_before, my_nt, sp.x );
            __syncwarp();
        }
        // candidates of piece `pc` -> S.raw, asynchronously (per-lane cp.async: the loads of the
        // next piece are in flight while the current one is swept, at no register cost)
        auto prefetch_piece = [&]( int pc )
        {
            if constexpr ( !TMA )
            {
#pragma unroll
                for ( int j = 0; j < kPieceEntries / 32; ++j )
                {
                    const int e = (int)lane + 32 * j;
                    const int tl = min( pc * kPieceTiles + ( e >> 3 ), T - 1 );
                    cp_async16( &S.raw[e],
                                a.q + S.tsrc[tl & ( kTableTiles - 1 )] + (unsigned)( e & 7 ) );
                }
                cp_async_commit();
            }
        };
        prefetch_piece( 0 );

        // ---- pieces of 16 tiles; a mask chunk is two pieces ---------------------------
        const int npieces = ( T + kPieceTiles - 1 ) / kPieceTiles;
        int in_chunk = 0, chunk_i = 0;
        for ( int pc = 0; pc < npieces; ++pc )
        {
            const int buf = pc & 1;
            float4 r[kPieceEntries / 32];
            if constexpr ( TMA )
            {
                if ( pc + 1 < npieces )
                    issue_piece( S, a.q, pc + 1, buf ^ 1, lane );
                while ( !mbar_try_wait( &S.mbar[buf], ( phase >> buf ) & 1u ) )
                {
                }
                phase ^= 1u << buf;
#pragma unroll
                for ( int j = 0; j < kPieceEntries / 32; ++j )
                    r[j] = S.raw[buf][(int)lane + 32 * j];
            }
            else
            {
                cp_async_wait_all();
#pragma unroll
                for ( int j = 0; j < kPieceEntries / 32; ++j )
                    r[j] = S.raw[(int)lane + 32 * j]; // (each lane reads what it copied)
            }
            // transform: tile-local coordinates, tf32 hi/lo split, squared norm
#pragma unroll
            for ( int j = 0; j < kPieceEntries / 32; ++j )
            {
                const int e = (int)lane + 32 * j;
                store_features( S.feat, e, r[j].x - Ox, r[j].y - Oy, r[j].z - Oz );
                if constexpr ( !TMA )
                    S.cpid[e] = __float_as_int( r[j].w );
            }
            __syncwarp();
            if constexpr ( !TMA )
            {
                if ( pc + 1 < npieces )
                {
                    if ( ( ( pc + 1 ) * kPieceTiles ) % kTableTiles == 0 )
                    {
                        build_tile_table( S.tsrc, ( pc + 1 ) * kPieceTiles, tiles_before, my_nt,
                                          sp.x );
                        __syncwarp();
                    }
                    prefetch_piece( pc + 1 );
                }
            }
            const int nt_p = min( kPieceTiles, T - pc * kPieceTiles );
            const int n_filter = min( max( T0 - pc * kPieceTiles, 0 ), nt_p );
            if constexpr ( TMA )
                sweep_segment<HALF, DIAG, 4>( a, H, S.feat,
                                              reinterpret_cast<const int*>( &S.raw[buf][0] ) + 3,
                                              nt_p, n_filter, a.tau, lane );
            else
                sweep_segment<HALF, DIAG, 1>( a, H, S.feat, S.cpid, nt_p, n_filter, a.tau, lane );
            in_chunk += nt_p;
            if ( in_chunk == kChunkTiles || pc == npieces - 1 )
            {
                flush_chunk<DIAG>( a, H, in_chunk, chunk_i, (size_t)chunk0, lane );
                in_chunk = 0;
                ++chunk_i;
            }
            __syncwarp();
        }
        H.cnt_g += __shfl_xor_sync( kFullMask, H.cnt_g, 1 );
        H.cnt_g += __shfl_xor_sync( kFullMask, H.cnt_g, 2 );
        H.cnt_g8 += __shfl_xor_sync( kFullMask, H.cnt_g8, 1 );
        H.cnt_g8 += __shfl_xor_sync( kFullMask, H.cnt_g8, 2 );
        if ( !DIAG )
        {
            // counts[] in particle order (the list's array) and in sorted order (fill pass)
            const int t = (int)( lane & 3u ), g = (int)( lane >> 2 );
            if ( t == 0 && H.act_g )
            {
                a.counts[H.pid_g] = H.cnt_g;
                a.cnt_sorted[rc.first + g] = H.cnt_g;
            }
            if ( t == 1 && H.act_g8 )
            {
                a.counts[H.pid_g8] = H.cnt_g8;
                a.cnt_sorted[rc.first + g + 8] = H.cnt_g8;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// fill pass: masks -> rows, staged in shared memory, written once at their final place
//   CSR  neighbors[offsets[pid] + k]   (offsets = exclusive scan of counts in particle
//                                       order, Cabana_VerletList.hpp:478-491, :507-523)
//   2D   neighbors[pid*width + k]
// ---------------------------------------------------------------------------------------
struct __align__( 16 ) FillSmem
{
    int rows[kRowCap];
    unsigned ids[kChunkEntries]; // ids of the chunk's candidates, tiles in REVERSE order
    unsigned tsrc[kTableTiles];
};

// Expand one mask word: tile i of the chunk is bit 31 - i, and its ids sit at
// ids2[8 * (31 - i)], so the bit position indexes the table directly.
// bfind = position of the most significant set bit (one FLO; 31 - __clz costs two more adds).
CB_D unsigned top_bit( unsigned m )
{
    unsigned p;
    asm( "bfind.u32 %0, %1;" : "=r"( p ) : "r"( m ) );
    return p;
}
// The expansion loop is the hot spot of the fill pass (one iteration per stored neighbour and
// lane): 32-bit shared-window addresses keep it at FLO, LEA, LDS, SHF, LOP3, STS, IADD, BRA;
// the table and the output are given as shared-window addresses.
CB_D void expand_word_at( unsigned mm, unsigned tab, unsigned out )
{
    while ( mm )
    {
        const unsigned p = top_bit( mm );
        asm( "xor.b32 %0, %0, %1;" : "+r"( mm ) : "r"( 1u << p ) );
        unsigned v;
        asm volatile( "ld.shared.u32 %0, [%1];" : "=r"( v ) : "r"( tab + ( p << 5 ) ) );
        asm volatile( "st.shared.u32 [%0], %1;" ::"r"( out ), "r"( v ) : "memory" );
        out += 4u;
    }
}
CB_D void expand_word_global( unsigned mm, const unsigned* ids2, int*& out )
{
    while ( mm )
    {
        const unsigned p = top_bit( mm );
        mm ^= 1u << p;
        *out++ = (int)ids2[kTileCands * p];
    }
}

template <bool CSR>
__global__ void __launch_bounds__( kBlockT, 4 )
    k_tile_fill( const __grid_constant__ TileArgs a )
{
    extern __shared__ __align__( 16 ) unsigned char s_dyn[];
    const unsigned lane = threadIdx.x & 31u;
    const int wib = threadIdx.x >> 5;
    FillSmem& S = reinterpret_cast<FillSmem*>( s_dyn )[wib];
    const int g = (int)( lane >> 2 ), t = (int)( lane & 3u );
    if ( a.spec_total && ( *a.spec_total > a.spec_capacity || *a.overflow != 0 ) )
    {
        if ( blockIdx.x == 0 && threadIdx.x == 0 )
            *a.spec_failed = 1; // the host sizes the buffers and launches the pass again
        return;
    }
    const int ntiles = *a.ntiles_dev;

    // The kernel is bound by the latency of dependent global loads, so everything a tile needs
    // is requested as early as its address is known: tickets two tiles ahead, the record of
    // the next tile during the current one, row bookkeeping in sorted order (no gather through
    // the permutation), masks and candidate ids one chunk ahead.
    auto grab = [&]() -> unsigned
    {
        unsigned v = 0u;
        if ( lane == 0u )
            v = atomicAdd( a.ticket + 1, 1u );
        return __shfl_sync( kFullMask, v, 0 );
    };
    unsigned tile = grab();
    unsigned tile_next = grab();
    unsigned tile_nn = 0u;
    uint4 rec_raw = make_uint4( 0u, 0u, 0u, 0u ), rec_next = rec_raw;
    int chunk0 = 0, chunk0_next = 0;
    if ( tile < (unsigned)ntiles )
    {
        rec_raw = a.recs[tile];
        chunk0 = a.chunk_off[tile];
    }
    for ( ;; tile = tile_next, tile_next = tile_nn, rec_raw = rec_next, chunk0 = chunk0_next )
    {
        if ( tile >= (unsigned)ntiles )
            break;
        tile_nn = grab();
        if ( tile_next < (unsigned)ntiles )
        {
            rec_next = a.recs[tile_next];
            chunk0_next = a.chunk_off[tile_next];
        }
        const TileRec rc = unpack_rec( rec_raw );
        uint4 m_next = a.masks[(size_t)chunk0 * 32u + lane];
        int pid = -1, cnt = 0;
        long long dst = 0;
        if ( (int)lane < rc.np )
        {
            cnt = a.cnt_sorted[rc.first + lane];
            if ( CSR )
                dst = (long long)a.dst_sorted[rc.first + lane];
            else
                pid = (int)a.permute[rc.first + lane];
        }
        uint2 sp = make_uint2( 0u, 0u );
        if ( lane < 9u )
            sp = a.spans[(size_t)tile * 9u + lane];
        if ( !CSR )
            dst = (long long)pid * a.width;
        // inclusive scans: the 16 row sizes; the mma tiles of the 9 windows
        int inc = cnt;
        const int my_nt = (int)( sp.y / kTileCands );
        int incl = my_nt;
#pragma unroll
        for ( int o = 1; o < 16; o <<= 1 )
        {
            const int y = __shfl_up_sync( kFullMask, inc, o );
            const int y2 = __shfl_up_sync( kFullMask, incl, o );
            if ( (int)lane >= o )
            {
                inc += y;
                incl += y2;
            }
        }
        const int total = __shfl_sync( kFullMask, inc, 15 );
        if ( total == 0 )
            continue;
        const int T = __shfl_sync( kFullMask, incl, 15 );
        const int tiles_before = incl - my_nt;
        __syncwarp();
        build_tile_table( S.tsrc, 0, tiles_before, my_nt, sp.x );
        int table_seg = 0;
        __syncwarp();
        const int nchunks = ( T + kChunkTiles - 1 ) / kChunkTiles;
        // ids of the candidates of chunk ci, one register per 32 entries (pad slots carry -1;
        // their mask bits are zero)
        // (lane i takes mma tile i of the chunk: its 8 ids are 32 contiguous, aligned bytes)
        uint4 idr[2];
        auto load_ids = [&]( int ci )
        {
            const int seg = ( ci * kChunkTiles ) / kTableTiles;
            if ( seg != table_seg )
            {
                __syncwarp();
                build_tile_table( S.tsrc, seg * kTableTiles, tiles_before, my_nt, sp.x );
                table_seg = seg;
                __syncwarp();
            }
            const int tl = min( ci * kChunkTiles + (int)lane, T - 1 );
            const uint4* src =
                reinterpret_cast<const uint4*>( a.permute + S.tsrc[tl & ( kTableTiles - 1 )] );
            idr[0] = src[0];
            idr[1] = src[1];
        };

        // Row windows [ha, hb): as many rows as fit the staging buffer (normally all 16);
        // a single row longer than the buffer goes straight to global memory.
        int ha = 0;
        while ( ha < kTileHomes )
        {
            const int base = ha == 0 ? 0 : __shfl_sync( kFullMask, inc, ha - 1 );
            const unsigned fits = __ballot_sync(
                kFullMask, (int)lane >= ha && lane < 16u && inc - base <= kRowCap );
            // inc is non-decreasing: the rows that fit are a run of ones starting at ha
            int hb = ha + ( __ffs( (int)~( fits >> ha ) ) - 1 );
            bool direct = false;
            if ( hb == ha )
            {
                hb = ha + 1;
                direct = true;
            }
            const int rowstart = inc - cnt - base; // valid for lanes in [ha, hb)
            const bool in_g = g >= ha && g < hb;
            const bool in_g8 = g + 8 >= ha && g + 8 < hb;
            const int wtotal = __shfl_sync( kFullMask, inc, hb - 1 ) - base;
            if ( wtotal > 0 )
            {
                int cur_g = __shfl_sync( kFullMask, rowstart, g );
                int cur_g8 = __shfl_sync( kFullMask, rowstart, g + 8 );
                int* gout_g = a.neighbors + __shfl_sync( kFullMask, dst, g );
                int* gout_g8 = a.neighbors + __shfl_sync( kFullMask, dst, g + 8 );
                if ( ha != 0 )
                    m_next = a.masks[(size_t)chunk0 * 32u + lane];
                load_ids( 0 );
                for ( int ci = 0; ci < nchunks; ++ci )
                {
                    uint4 m = m_next;
                    __syncwarp(); // the previous chunk's walk is done with S.ids
                    {
                        uint4* d = reinterpret_cast<uint4*>( S.ids + ( 31 - (int)lane ) * kTileCands );
                        d[0] = idr[0];
                        d[1] = idr[1];
                    }
                    if ( ci + 1 < nchunks )
                    {
                        m_next = a.masks[( (size_t)chunk0 + ci + 1 ) * 32u + lane];
                        load_ids( ci + 1 );
                    }
                    __syncwarp();
                    if ( !in_g )
                        m.x = m.y = 0u;
                    if ( !in_g8 )
                        m.z = m.w = 0u;
                    // where this thread's hits go: exclusive prefix over the quad
                    const int pg = __popc( m.x ) + __popc( m.y );
                    const int pg8 = __popc( m.z ) + __popc( m.w );
                    int pk = pg | ( pg8 << 16 );
                    int y = __shfl_up_sync( kFullMask, pk, 1, 4 );
                    if ( t >= 1 )
                        pk += y;
                    y = __shfl_up_sync( kFullMask, pk, 2, 4 );
                    if ( t >= 2 )
                        pk += y;
                    const int qt = __shfl_sync( kFullMask, pk, 3, 4 );
                    const unsigned* ids2 = S.ids + 2 * t;
                    if ( !direct )
                    {
                        // The four words of a lane are expanded by four loops the whole warp
                        // walks together, each as long as its busiest lane.  Every word knows
                        // where its ids go, so a lane may take them in any order: sorted by
                        // population (largest first), the first loop is the only long one.
                        const int w_g = cur_g + ( pk & 0xffff ) - pg;
                        const int w_g8 = cur_g8 + ( pk >> 16 ) - pg8;
                        const unsigned rows_a = smem_u32( S.rows );
                        const unsigned tab0 = smem_u32( ids2 );
                        unsigned wm[4] = { m.x, m.y, m.z, m.w };
                        unsigned wo[4] = { rows_a + 4u * (unsigned)w_g,
                                           rows_a + 4u * (unsigned)( w_g + __popc( m.x ) ),
                                           rows_a + 4u * (unsigned)w_g8,
                                           rows_a + 4u * (unsigned)( w_g8 + __popc( m.z ) ) };
                        unsigned wt[4] = { tab0, tab0 + 4u, tab0, tab0 + 4u };
                        int wc[4] = { __popc( m.x ), __popc( m.y ), __popc( m.z ), __popc( m.w ) };
                        auto cswap = [&]( int i, int j )
                        {
                            const bool sw = wc[i] < wc[j];
                            const unsigned tm = sw ? wm[j] : wm[i], to = sw ? wo[j] : wo[i],
                                           tt = sw ? wt[j] : wt[i];
                            const int tc = sw ? wc[j] : wc[i];
                            wm[j] = sw ? wm[i] : wm[j];
                            wo[j] = sw ? wo[i] : wo[j];
                            wt[j] = sw ? wt[i] : wt[j];
                            wc[j] = sw ? wc[i] : wc[j];
                            wm[i] = tm;
                            wo[i] = to;
                            wt[i] = tt;
                            wc[i] = tc;
                        };
                        if ( a.fill_sort )
                        {
                            cswap( 0, 1 );
                            cswap( 2, 3 );
                            cswap( 0, 2 );
                            cswap( 1, 3 );
                            cswap( 1, 2 );
                        }
#pragma unroll
                        for ( int k = 0; k < 4; ++k )
                            expand_word_at( wm[k], wt[k], wo[k] );
                        cur_g += qt & 0xffff;
                        cur_g8 += qt >> 16;
                    }
                    else
                    {
                        int* w_g = gout_g + ( ( pk & 0xffff ) - pg );
                        int* w_g8 = gout_g8 + ( ( pk >> 16 ) - pg8 );
                        expand_word_global( m.x, ids2, w_g );
                        expand_word_global( m.y, ids2 + 1, w_g );
                        expand_word_global( m.z, ids2, w_g8 );
                        expand_word_global( m.w, ids2 + 1, w_g8 );
                        gout_g += qt & 0xffff;
                        gout_g8 += qt >> 16;
                    }
                }
                __syncwarp();
                if ( !direct )
                {
                    // rows leave shared memory coalesced, each to its final place; up to
                    // 96 ids per row without a loop
#pragma unroll 1
                    for ( int h = ha; h < hb; ++h )
                    {
                        const int c = __shfl_sync( kFullMask, cnt, h );
                        const int rs = __shfl_sync( kFullMask, rowstart, h );
                        int* d = a.neighbors + __shfl_sync( kFullMask, dst, h ) + lane;
                        const int* src = S.rows + rs + lane;
                        const int left = c - (int)lane;
                        if ( left > 0 )
                            __stcs( d, src[0] );
                        if ( left > 32 )
                            __stcs( d + 32, src[32] );
                        if ( left > 64 )
                            __stcs( d + 64, src[64] );
#pragma unroll 1
                        for ( int i = 96; i < left; i += 32 )
                            __stcs( d + i, src[i] );
                    }
                }
                __syncwarp();
            }
            ha = hb;
        }
    }
}

// dst_sorted[s] = offsets[permute[s]]: the CSR row start of the particle in sorted slot s, so
// the fill pass reads its row bookkeeping coalesced instead of through the permutation.
__global__ void __launch_bounds__( 256 )
    k_sorted_dst( const unsigned* __restrict__ permute, const int* __restrict__ offsets,
                  const unsigned* __restrict__ cell_off, long long ncells,
                  int* __restrict__ dst_sorted )
{
    const long long ns = cell_off[ncells];
    for ( long long s = (long long)blockIdx.x * 256 + threadIdx.x; s < ns;
          s += (long long)gridDim.x * 256 )
    {
        const unsigned pid = permute[s];
        dst_sorted[s] = pid == 0xffffffffu ? 0 : offsets[pid];
    }
}

int persistent_blocks( const void* func, int threads, int smem )
{
    int nb = 0;
    if ( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &nb, func, threads, smem ) !=
             cudaSuccess ||
         nb < 1 )
    {
        (void)cudaGetLastError();
        nb = 1;
    }
    int dev = 0, sms = kNumSMs;
    if ( cudaGetDevice( &dev ) == cudaSuccess )
        cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, dev );
    return nb * sms;
}

template <class K>
int launch_persistent( K kernel, int smem, const TileArgs& a, cudaStream_t stream,
                       int threads = kBlockT )
{
    // per device/context: set unconditionally (cheap) rather than caching per process
    CB_CUDA( cudaFuncSetAttribute( kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   smem ) );
    const int blocks = persistent_blocks( (const void*)kernel, threads, smem );
    kernel<<<blocks, threads, smem, stream>>>( a );
    CB_CHECK_LAUNCH();
    return CB_OK;
}

} // namespace

// =======================================================================================
// host side
// =======================================================================================
void make_tile_grid( TileGrid& tg, const double* grid_min, const double* grid_max,
                     double radius, long long n )
{
    memset( &tg, 0, sizeof( tg ) );
    const double reach = radius * ( 1.0 + 1.0e-6 );
    int nc[3];
    for ( int d = 0; d < 3; ++d )
    {
        tg.g.min[d] = grid_min[d];
        tg.g.max[d] = grid_max[d];
    }
    for ( int d = 0; d < 2; ++d )
    {
        const double L = grid_max[d] - grid_min[d];
        double c = floor( L / reach );
        if ( !( c >= 1.0 ) )
            c = 1.0;
        if ( c > 30000.0 )
            c = 30000.0; // wider columns are still correct (every column >= reach)
        nc[d] = (int)c;
    }
    // keep the column count sane relative to the particle count (sparse boxes)
    while ( (double)nc[0] * nc[1] > 4.0 * (double)n + 4096.0 )
    {
        if ( nc[0] >= nc[1] && nc[0] > 1 )
            nc[0] = ( nc[0] + 1 ) / 2;
        else if ( nc[1] > 1 )
            nc[1] = ( nc[1] + 1 ) / 2;
        else
            break;
    }
    const double Lz = grid_max[2] - grid_min[2];
    const char* div_env = getenv( "CB_TILE_ZDIV" );
    double zdiv = div_env ? atof( div_env ) : 4.0;
    if ( !( zdiv >= 1.0 ) || zdiv > 16.0 )
        zdiv = 4.0;
    double cz = floor( Lz / ( radius / zdiv ) );
    if ( !( cz >= 1.0 ) )
        cz = 1.0;
    if ( cz > 32767.0 )
        cz = 32767.0;
    nc[2] = (int)cz;
    const double cell_limit = 8.0 * (double)n + 1048576.0;
    while ( nc[2] > 1 && (double)nc[0] * nc[1] * nc[2] > cell_limit )
        nc[2] = ( nc[2] + 1 ) / 2;
    for ( int d = 0; d < 3; ++d )
    {
        tg.g.nx[d] = nc[d];
        tg.g.dx[d] = ( grid_max[d] - grid_min[d] ) / (double)nc[d];
        tg.g.rdx[d] = 1.0 / tg.g.dx[d];
    }
    tg.ncx = nc[0];
    tg.ncy = nc[1];
    tg.nz = nc[2];
    // |z_i - z_j| <= r  =>  cells differ by at most floor(r/h) + 1
    tg.kz = (int)floor( reach / tg.g.dx[2] ) + 1;
    const char* zb_env = getenv( "CB_TILE_ZB" );
    tg.zb = zb_env ? atoi( zb_env ) : 16;
    if ( tg.zb < 1 || tg.zb > 64 )
        tg.zb = 16;
    tg.nzb = ( tg.nz + tg.zb - 1 ) / tg.zb;
    tg.ncols = (long long)tg.ncx * tg.ncy;
    while ( tg.ncols * tg.nzb >= ( 1ll << 30 ) )
    {
        tg.zb *= 2;
        tg.nzb = ( tg.nz + tg.zb - 1 ) / tg.zb;
    }
    tg.nblocks = tg.ncols * tg.nzb;
    tg.ncells = tg.ncols * tg.nz;
}

double tile_filter_bound( const TileGrid& tg, double radius, int nzc )
{
    // See DESIGN.md "Exactness (v2)".  u = 2^-24 (fp32 unit roundoff), v = 2^-20 (what the
    // hi/lo tf32 split drops), gam = 2^-21 per accumulated sum inside the tensor core.
    // nzc = z cells the home particles of one staging span (tile-local origin at their centre).
    const double u = ldexp( 1.0, -24 ), v = ldexp( 1.0, -20 ), gam = ldexp( 1.0, -21 );
    double M = 0.0;
    for ( int d = 0; d < 3; ++d )
        M = fmax( M, tg.g.max[d] - tg.g.min[d] );
    M *= 1.0 + 1.0e-9;
    const double sl = 1.0 + 1.0e-5;
    const double Dx = 1.5 * tg.g.dx[0] * sl, Dy = 1.5 * tg.g.dx[1] * sl;
    const double Dz = ( 0.5 * nzc + tg.kz + 1.0 ) * tg.g.dx[2] * sl;
    const double Hx = 0.5 * tg.g.dx[0] * sl, Hy = 0.5 * tg.g.dx[1] * sl;
    const double Hz = ( 0.5 * nzc + 0.5 ) * tg.g.dx[2] * sl;
    const double Dmax = fmax( Dx, fmax( Dy, Dz ) );
    const double S = Dx * Dx + Dy * Dy + Dz * Dz;
    const double Sh = Hx * Hx + Hy * Hy + Hz * Hz;
    const double HD = Hx * Dx + Hy * Dy + Hz * Dz;
    const double rsqr = radius * radius;
    // (1)+(2) coordinate roundings (q = fl32(x - min), x' = fl32(q - O)), for s <= 4 r^2
    double E = 4.0 * sqrt( 3.0 ) * u * ( M + Dmax ) * ( 2.02 * radius ) +
               12.0 * u * u * ( M * M + Dmax * Dmax );
    // (3) norms in fp32, hi/lo splits, cutoff split, accumulation
    E += 3.0 * u * ( S + Sh ) + v * ( S + Sh ) + 6.0 * v * HD + v * rsqr;
    E += 2.0 * u * ( Sh + rsqr ); // accumulator init fl(fl(N - r2hi) - r2lo)
    E += 2.0 * gam * ( 2.0 * HD * ( 1.0 + 1.0e-3 ) + S + Sh + rsqr );
    return E;
}

int tile_bin( const TileGrid& tg, const cb_positions& x, int* cell_counts, unsigned* cell_off,
              uint2* cellslot, unsigned char* pads, float4* q, unsigned* permute,
              DeviceBuffer& scan_scratch, cudaStream_t stream, const long long* n_dev )
{
    const long long n = x.n;
    CB_CUDA( cudaMemsetAsync( cell_counts, 0, sizeof( int ) * (size_t)tg.ncells, stream ) );
    if ( n > 0 )
    {
        k_tbin_count<<<launch_grid_for( n, 256 * kBinBatch ), 256, 0, stream>>>(
            make_access( x ), to_grid( tg.g ), n, cell_counts, cellslot, n_dev );
        CB_CHECK_LAUNCH();
    }
    k_tbin_pad<<<launch_grid_for( tg.ncols * 32, 256 ), 256, 0, stream>>>( cell_counts, tg.ncols,
                                                                           tg.nz, pads );
    CB_CHECK_LAUNCH();
    CB_TRY( exclusive_scan_i32( cell_counts, reinterpret_cast<int*>( cell_off ), tg.ncells,
                                true, nullptr, scan_scratch, stream ) );
    if ( n > 0 )
    {
        k_tbin_scatter<<<launch_grid_for( n, 256 * kBinBatch ), 256, 0, stream>>>(
            make_access( x ), n, cellslot, cell_off, q, permute, tg.g.min[0], tg.g.min[1],
            tg.g.min[2], n_dev );
        CB_CHECK_LAUNCH();
    }
    k_tbin_sentinels<<<launch_grid_for( tg.ncols + 1, 256 ), 256, 0, stream>>>(
        cell_off, pads, tg.ncols, tg.nz, q, permute );
    CB_CHECK_LAUNCH();
    return CB_OK;
}

int tile_plan( const TileGrid& tg, const unsigned* cell_off, bool half, int* block_tiles,
               int* tile_base, uint4* recs, uint2* spans, int* tile_chunks, int* chunk_off,
               long long rec_capacity, DeviceBuffer& scan_scratch, cudaStream_t stream )
{
    const GridInts gi = { tg.ncx, tg.ncy, tg.nz, tg.kz };
    k_plan_blocks<<<launch_grid_for( tg.nblocks, 256 ), 256, 0, stream>>>(
        cell_off, tg.nz, tg.zb, tg.nzb, tg.nblocks, block_tiles );
    CB_CHECK_LAUNCH();
    CB_TRY( exclusive_scan_i32( block_tiles, tile_base, tg.nblocks, true, nullptr,
                                scan_scratch, stream ) );
    CB_CUDA( cudaMemsetAsync( tile_chunks, 0, sizeof( int ) * (size_t)rec_capacity, stream ) );
    k_plan_tiles<<<launch_grid_for( tg.nblocks * kPlanSplit, 128 ), 128, 0, stream>>>(
        cell_off, gi, tg.zb, tg.nzb, tg.nblocks, half ? 1 : 0, block_tiles, tile_base, recs,
        spans, tile_chunks, rec_capacity );
    CB_CHECK_LAUNCH();
    CB_TRY( exclusive_scan_i32( tile_chunks, chunk_off, rec_capacity, true, nullptr,
                                scan_scratch, stream ) );
    return CB_OK;
}

// CB_TILE_STAGING = tma (default: cp.async.bulk of every contiguous candidate span, mbarrier
// completion, double buffered) | async (per-lane cp.async gather through a tile table).
static bool use_tma_staging()
{
    const char* e = getenv( "CB_TILE_STAGING" );
    return !( e && strcmp( e, "async" ) == 0 );
}

template <bool DIAG>
static int launch_count( const TileArgs& a, bool half, cudaStream_t stream )
{
    if ( use_tma_staging() )
    {
        const int smem = (int)sizeof( CountSmemTma ) * kWarpsT;
        return half ? launch_persistent( k_tile_count<true, DIAG, true>, smem, a, stream )
                    : launch_persistent( k_tile_count<false, DIAG, true>, smem, a, stream );
    }
    const int smem = (int)sizeof( CountSmemLdg ) * kWarpsT;
    return half ? launch_persistent( k_tile_count<true, DIAG, false>, smem, a, stream )
                : launch_persistent( k_tile_count<false, DIAG, false>, smem, a, stream );
}

int tile_count_pass( const TileArgs& a, bool half, cudaStream_t stream )
{
    return launch_count<false>( a, half, stream );
}

int tile_diag_pass( const TileArgs& a, bool half, cudaStream_t stream )
{
    return launch_count<true>( a, half, stream );
}

int tile_sorted_dst( const TileArgs& a, long long ncells, long long ns_cap, int* dst_sorted,
                     cudaStream_t stream )
{
    k_sorted_dst<<<launch_grid_for( ns_cap, 256 ), 256, 0, stream>>>( a.permute, a.offsets,
                                                                     a.cell_off, ncells, dst_sorted );
    CB_CHECK_LAUNCH();
    return CB_OK;
}

int tile_fill_pass( const TileArgs& a_, bool csr, cudaStream_t stream )
{
    TileArgs a = a_;
    const char* sort_env = getenv( "CB_FILL_SORT" );
    a.fill_sort = !( sort_env && sort_env[0] == '0' );
    const int smem = (int)sizeof( FillSmem ) * kWarpsT;
    return csr ? launch_persistent( k_tile_fill<true>, smem, a, stream )
               : launch_persistent( k_tile_fill<false>, smem, a, stream );
}

} // namespace cb
