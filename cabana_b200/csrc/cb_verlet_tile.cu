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
constexpr int kRowCap = 1536;                           // ids staged per tile in the fill pass

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
    start = cell_off[c0 + z0];
    len = cell_off[c0 + z1 + 1] - start;
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
// gather: the FP32 origin-relative copy of the positions in cell-sorted order
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__( 256 )
    k_gather_q( PosAccess x, long long n, const unsigned* __restrict__ permute,
                float4* __restrict__ q, double ox, double oy, double oz )
{
    for ( long long s = (long long)blockIdx.x * 256 + threadIdx.x; s < n + 8;
          s += (long long)gridDim.x * 256 )
    {
        if ( s >= n )
        {
            q[s] = make_float4( 1.0e18f, 1.0e18f, 1.0e18f, __int_as_float( -1 ) );
            continue;
        }
        const unsigned pid = permute[s];
        const long long off = x.offset( (long long)pid );
        const double px = x.base[off];
        const double py = x.base[off + x.comp_stride];
        const double pz = x.base[off + 2 * x.comp_stride];
        q[s] = make_float4( __double2float_rn( px - ox ), __double2float_rn( py - oy ),
                            __double2float_rn( pz - oz ), __int_as_float( (int)pid ) );
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

__global__ void __launch_bounds__( 128 )
    k_plan_tiles( const unsigned* __restrict__ cell_off, GridInts gi, int zb, int nzb,
                  long long nblocks, int half, const int* __restrict__ block_tiles,
                  const int* __restrict__ tile_base, uint4* __restrict__ recs,
                  uint2* __restrict__ spans, int* __restrict__ tile_chunks,
                  long long rec_capacity )
{
    for ( long long b = (long long)blockIdx.x * 128 + threadIdx.x; b < nblocks;
          b += (long long)gridDim.x * 128 )
    {
        const int nt = block_tiles[b];
        if ( nt == 0 )
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
        for ( int ti = 0; ti < nt; ++ti )
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

// Shared memory of one warp.  TMA staging: candidates land in `raw` by cp.async.bulk.
// LDG staging: every lane fetches its candidates itself through a per-tile source table.
struct __align__( 16 ) CountSmemTma
{
    float4 f0[kPieceEntries];     // (hx, hy, hz, n_hi) of the staged candidates
    float4 f1[kPieceEntries];     // (lx, ly, lz, n_lo)
    float4 raw[2][kPieceEntries]; // bulk-copy landing zone (q records), double buffered
    unsigned sp_start[16], sp_len[16], sp_pos[16];
    float kscr[16]; // |x_i|^2 - r^2 of the home particles (accumulator init)
    unsigned long long mbar[2];
    CB_D int cand_pid( int buf, int e ) const { return __float_as_int( raw[buf][e].w ); }
};
struct __align__( 16 ) CountSmemLdg
{
    float4 f0[kPieceEntries];
    float4 f1[kPieceEntries];
    int cpid[kPieceEntries];        // particle ids of the staged candidates (exact tier)
    unsigned tsrc[kTableTiles];     // sorted slot of the first entry of every mma tile
    unsigned char tval[kTableTiles]; // valid entries (1..8) of every mma tile
    float kscr[16];
    CB_D int cand_pid( int, int e ) const { return cpid[e]; }
};

// Stage piece `pc` (candidate list entries [128 pc, 128 pc + 128)) into raw[buf]: every span
// that intersects the piece is one bulk copy (16-byte records, 16-byte aligned on both
// sides); the <= 7 padding entries behind a span's end and the tail behind the last tile
// are filled with far-away sentinels by ordinary stores (disjoint from what the copies
// write, so no proxy fence is needed; the buffer's previous readers finished before the
// __syncwarp that precedes this call).  All lanes call.
CB_D void issue_piece( CountSmemTma& S, const float4* __restrict__ q, int pc, int buf,
                       int total_entries, unsigned lane )
{
    const float4 sent = make_float4( 1.0e18f, 1.0e18f, 1.0e18f, __int_as_float( -1 ) );
    const unsigned p0 = (unsigned)( pc * kPieceEntries );
    const unsigned p1 = p0 + kPieceEntries;
    unsigned lo = 0u, hi = 0u, p = 0u;
    if ( lane < 9u )
    {
        p = S.sp_pos[lane];
        const unsigned end = p + S.sp_len[lane];
        lo = max( p, p0 );
        hi = min( end, p1 );
        if ( hi < lo )
            hi = lo;
        // padding behind this span's end, if the end lies in this piece
        if ( end > p0 && end <= p1 && S.sp_len[lane] != 0u )
            for ( unsigned e = end; ( e & 7u ) != 0u; ++e )
                S.raw[buf][e - p0] = sent;
    }
    // tail behind the last tile (last piece only)
    for ( unsigned e = max( (unsigned)total_entries, p0 ) + lane; e < p1; e += 32u )
        S.raw[buf][e - p0] = sent;
    const unsigned bytes = ( hi - lo ) * 16u;
    const unsigned total = __reduce_add_sync( kFullMask, bytes );
    if ( lane == 0u )
        mbar_expect_tx( &S.mbar[buf], total );
    __syncwarp();
    if ( bytes )
        bulk_g2s( &S.raw[buf][lo - p0], q + S.sp_start[lane] + ( lo - p ), bytes,
                  &S.mbar[buf] );
}

// Source table of the mma tiles [seg0, seg0 + kTableTiles): the lane that owns span s
// writes the entries of its tiles (tile k of the span starts at sorted slot start + 8k).
template <class SM>
CB_D void build_tile_table( SM& S, int seg0, int tiles_before, int my_nt, unsigned sp_start,
                            unsigned sp_len )
{
    const int i0 = max( 0, seg0 - tiles_before );
    const int i1 = min( my_nt, seg0 + kTableTiles - tiles_before );
    for ( int i = i0; i < i1; ++i )
    {
        const int k = tiles_before + i - seg0;
        S.tsrc[k] = sp_start + (unsigned)( kTileCands * i );
        S.tval[k] = (unsigned char)min( (unsigned)kTileCands,
                                        sp_len - (unsigned)( kTileCands * i ) );
    }
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
    const int g = (int)( lane >> 2 ), t = (int)( lane & 3u );
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
    float* f0w = reinterpret_cast<float*>( S.f0 );
    float* f1w = reinterpret_cast<float*>( S.f1 );
    const float4 sent = make_float4( 1.0e18f, 1.0e18f, 1.0e18f, __int_as_float( -1 ) );

    for ( ;; )
    {
        unsigned tile = 0u;
        if ( lane == 0u )
            tile = atomicAdd( a.ticket, 1u );
        tile = __shfl_sync( kFullMask, tile, 0 );
        if ( tile >= (unsigned)ntiles )
            break;
        const TileRec rc = unpack_rec( a.recs[tile] );
        const int chunk0 = a.chunk_off[tile];

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
        const float xh = hq.x - Ox, yh = hq.y - Oy, zh = hq.z - Oz;
        // candidate spans (planned): lane s < 9 owns span s
        uint2 sp = make_uint2( 0u, 0u );
        if ( lane < 9u )
            sp = a.spans[(size_t)tile * 9u + lane];
        __syncwarp();
        if ( lane < 16u )
        {
            // A operands, one row per home particle (see mma_tf32):
            //   A1 = [-2H, 1 | -2H, 1]   A2 = [-2L, 0 | -2L, 0]   C = |x_i|^2 - r^2
            const float Hx = trunc_tf32( xh ), Hy = trunc_tf32( yh ), Hz = trunc_tf32( zh );
            const float N = fmaf( zh, zh, fmaf( yh, yh, xh * xh ) );
            S.f0[lane] = make_float4( -2.f * Hx, -2.f * Hy, -2.f * Hz, 1.f );
            S.f1[lane] = make_float4( -2.f * ( xh - Hx ), -2.f * ( yh - Hy ),
                                      -2.f * ( zh - Hz ), 0.f );
            // (N - r2hi) - r2lo: N and r2hi are close in magnitude, r2lo is the small part
            S.kscr[lane] = ( N - a.r2hi ) - a.r2lo;
        }
        const int my_nt = (int)( ( sp.y + kTileCands - 1 ) / kTileCands );
        int incl = my_nt;
#pragma unroll
        for ( int o = 1; o < 16; o <<= 1 )
        {
            const int y = __shfl_up_sync( kFullMask, incl, o );
            if ( (int)lane >= o )
                incl += y;
        }
        const int T = __shfl_sync( kFullMask, incl, 15 );
        const int T0 = __shfl_sync( kFullMask, incl, 2 ); // tiles of the da = 0 spans
        const int tiles_before = incl - my_nt;
        // list position of home particle h: selfbase + h (the home span is s = 1)
        const unsigned selfbase =
            __shfl_sync( kFullMask, (unsigned)tiles_before * kTileCands - sp.x, 1 ) + rc.first;
        __syncwarp();
        const float a1_lo = f0w[4 * g + t], a1_hi = f0w[4 * ( g + 8 ) + t];
        const float a2_lo = f1w[4 * g + t], a2_hi = f1w[4 * ( g + 8 ) + t];
        const float k_g = S.kscr[g], k_g8 = S.kscr[g + 8];
        const float cinit[4] = { k_g, k_g, k_g8, k_g8 };
        const float hx_g = __shfl_sync( kFullMask, xh, g );
        const float hx_g8 = __shfl_sync( kFullMask, xh, g + 8 );
        const int pid_g = __shfl_sync( kFullMask, pid, g );
        const int pid_g8 = __shfl_sync( kFullMask, pid, g + 8 );
        const bool act_g = ( actmask >> g ) & 1u;
        const bool act_g8 = ( actmask >> ( g + 8 ) ) & 1u;
        __syncwarp();
        if constexpr ( TMA )
        {
            if ( lane < 16u )
            {
                S.sp_start[lane] = sp.x;
                S.sp_len[lane] = sp.y;
                S.sp_pos[lane] = (unsigned)tiles_before * kTileCands;
            }
            __syncwarp();
            issue_piece( S, a.q, 0, 0, T * kTileCands, lane );
        }
        else
        {
            build_tile_table( S, 0, tiles_before, my_nt, sp.x, sp.y );
            __syncwarp();
        }

        // ---- pieces ----------------------------------------------------------------
        const int npieces = ( T + kPieceTiles - 1 ) / kPieceTiles;
        unsigned m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;
        float ma = 3.0e38f;
        int cnt_g = 0, cnt_g8 = 0;
        int in_chunk = 0, chunk_i = 0;
        for ( int pc = 0; pc < npieces; ++pc )
        {
            const int buf = pc & 1;
            if constexpr ( TMA )
            {
                if ( pc + 1 < npieces )
                    issue_piece( S, a.q, pc + 1, buf ^ 1, T * kTileCands, lane );
                while ( !mbar_try_wait( &S.mbar[buf], ( phase >> buf ) & 1u ) )
                {
                }
                phase ^= 1u << buf;
                __syncwarp(); // the sentinel stores of other lanes
            }
            else if ( pc > 0 && ( pc * kPieceTiles ) % kTableTiles == 0 )
            {
                build_tile_table( S, pc * kPieceTiles, tiles_before, my_nt, sp.x, sp.y );
                __syncwarp();
            }
            // transform: tile-local coordinates, tf32 hi/lo split, squared norm
            float4 r[kPieceEntries / 32];
#pragma unroll
            for ( int j = 0; j < kPieceEntries / 32; ++j )
            {
                const int e = (int)lane + 32 * j;
                if constexpr ( TMA )
                    r[j] = S.raw[buf][e];
                else
                {
                    const int tl = pc * kPieceTiles + ( e >> 3 );
                    const int k = tl & ( kTableTiles - 1 );
                    r[j] = sent;
                    if ( tl < T && ( e & 7 ) < (int)S.tval[k] )
                        r[j] = a.q[S.tsrc[k] + (unsigned)( e & 7 )];
                }
            }
#pragma unroll
            for ( int j = 0; j < kPieceEntries / 32; ++j )
            {
                const int e = (int)lane + 32 * j;
                const float x = r[j].x - Ox, y = r[j].y - Oy, z = r[j].z - Oz;
                const float hx = trunc_tf32( x ), hy = trunc_tf32( y ), hz = trunc_tf32( z );
                const float n = fmaf( z, z, fmaf( y, y, x * x ) );
                const float nh = trunc_tf32( n );
                S.f0[e] = make_float4( hx, hy, hz, nh );
                S.f1[e] = make_float4( x - hx, y - hy, z - hz, n - nh );
                if constexpr ( !TMA )
                    S.cpid[e] = __float_as_int( r[j].w );
            }
            __syncwarp();
            const int nt_p = min( kPieceTiles, T - pc * kPieceTiles );
#pragma unroll 4
            for ( int j = 0; j < nt_p; ++j )
            {
                const float b0 = f0w[32 * j + (int)lane];
                const float b1 = f1w[32 * j + (int)lane];
                float c1[4], c[4];
                mma_tf32( c1, a1_lo, a1_hi, a1_lo, a1_hi, b0, b1, cinit );
                mma_tf32( c, a2_lo, a2_hi, a2_lo, a2_hi, b0, b1, c1 );
                m0 = __funnelshift_l( __float_as_uint( c[0] ), m0, 1 );
                m1 = __funnelshift_l( __float_as_uint( c[1] ), m1, 1 );
                m2 = __funnelshift_l( __float_as_uint( c[2] ), m2, 1 );
                m3 = __funnelshift_l( __float_as_uint( c[3] ), m3, 1 );
                ma = fminf( fminf( ma, fabsf( c[0] ) ), fabsf( c[1] ) );
                ma = fminf( fminf( ma, fabsf( c[2] ) ), fabsf( c[3] ) );
            }
            // ---- exact tier: values inside the filter's error band --------------------
            if ( DIAG || __any_sync( kFullMask, ma <= a.tau ) )
            {
                for ( int j = 0; j < nt_p; ++j )
                {
                    const float b0 = f0w[32 * j + (int)lane];
                    const float b1 = f1w[32 * j + (int)lane];
                    float c1[4], c[4];
                    mma_tf32( c1, a1_lo, a1_hi, a1_lo, a1_hi, b0, b1, cinit );
                    mma_tf32( c, a2_lo, a2_hi, a2_lo, a2_hi, b0, b1, c1 );
                    const unsigned bit = 1u << ( nt_p - 1 - j );
#pragma unroll
                    for ( int k = 0; k < 4; ++k )
                    {
                        const int e = kTileCands * j + 2 * t + ( k & 1 );
                        const int hp = k < 2 ? pid_g : pid_g8;
                        if ( DIAG )
                        {
                            const int cp = S.cand_pid( buf, e );
                            if ( hp >= 0 && cp >= 0 )
                            {
                                const float err =
                                    (float)fabs( (double)c[k] - exact_c( a, hp, cp ) );
                                atomicMax( a.diag_maxerr, __float_as_uint( err ) );
                            }
                        }
                        if ( fabsf( c[k] ) <= a.tau )
                        {
                            const bool hit = exact_decide<HALF>( a, hp, S.cand_pid( buf, e ) );
                            unsigned& m = k == 0 ? m0 : ( k == 1 ? m1 : ( k == 2 ? m2 : m3 ) );
                            m = hit ? ( m | bit ) : ( m & ~bit );
                        }
                    }
                }
            }
            ma = 3.0e38f;
            // ---- half lists: x-major criterion on the spans of the home x column ------
            if ( HALF && pc * kPieceTiles < T0 )
            {
                const unsigned low = ( 1u << nt_p ) - 1u; // nt_p <= 16
#pragma unroll
                for ( int k = 0; k < 4; ++k )
                {
                    unsigned& m = k == 0 ? m0 : ( k == 1 ? m1 : ( k == 2 ? m2 : m3 ) );
                    const float hx = k < 2 ? hx_g : hx_g8;
                    const int hp = k < 2 ? pid_g : pid_g8;
                    unsigned mm = m & low;
                    while ( mm )
                    {
                        const int b = 31 - __clz( mm );
                        mm &= ~( 1u << b );
                        const int j = nt_p - 1 - b;
                        if ( pc * kPieceTiles + j >= T0 )
                            continue; // da = +1: a larger cell index means a larger x
                        const int e = kTileCands * j + 2 * t + ( k & 1 );
                        const float cx = S.f0[e].x + S.f1[e].x; // == x - Ox exactly
                        bool keep;
                        if ( cx > hx )
                            keep = true; // float rounding is monotone
                        else if ( cx < hx )
                            keep = false;
                        else
                            keep = exact_decide<HALF>( a, hp, S.cand_pid( buf, e ) );
                        if ( !keep )
                            m &= ~( 1u << b );
                    }
                }
            }
            in_chunk += nt_p;
            if ( in_chunk == kChunkTiles || pc == npieces - 1 )
            {
                // normalise: tile i of the chunk sits at bit 31 - i
                if ( in_chunk < kChunkTiles )
                {
                    const int sh = kChunkTiles - in_chunk;
                    m0 <<= sh;
                    m1 <<= sh;
                    m2 <<= sh;
                    m3 <<= sh;
                }
                // j != i: the home particles sit at list entries selfbase .. selfbase + 15
                if ( (int)( selfbase / kChunkEntries ) <= chunk_i &&
                     (int)( ( selfbase + 15u ) / kChunkEntries ) >= chunk_i )
                {
#pragma unroll
                    for ( int hh = 0; hh < 2; ++hh )
                    {
                        const unsigned p = selfbase + (unsigned)( g + 8 * hh );
                        const unsigned e = p % kChunkEntries;
                        if ( (int)( p / kChunkEntries ) == chunk_i &&
                             (int)( ( e & 7u ) >> 1 ) == t )
                        {
                            const unsigned keep = ~( 0x80000000u >> ( e >> 3 ) );
                            if ( hh == 0 )
                            {
                                if ( e & 1u )
                                    m1 &= keep;
                                else
                                    m0 &= keep;
                            }
                            else
                            {
                                if ( e & 1u )
                                    m3 &= keep;
                                else
                                    m2 &= keep;
                            }
                        }
                    }
                }
                if ( !act_g )
                    m0 = m1 = 0u;
                if ( !act_g8 )
                    m2 = m3 = 0u;
                cnt_g += __popc( m0 ) + __popc( m1 );
                cnt_g8 += __popc( m2 ) + __popc( m3 );
                if ( !DIAG )
                    a.masks[( (size_t)chunk0 + chunk_i ) * 32u + lane] =
                        make_uint4( m0, m1, m2, m3 );
                m0 = m1 = m2 = m3 = 0u;
                in_chunk = 0;
                ++chunk_i;
            }
            __syncwarp();
        }
        cnt_g += __shfl_xor_sync( kFullMask, cnt_g, 1 );
        cnt_g += __shfl_xor_sync( kFullMask, cnt_g, 2 );
        cnt_g8 += __shfl_xor_sync( kFullMask, cnt_g8, 1 );
        cnt_g8 += __shfl_xor_sync( kFullMask, cnt_g8, 2 );
        if ( !DIAG )
        {
            if ( t == 0 && act_g )
                a.counts[pid_g] = cnt_g;
            if ( t == 1 && act_g8 )
                a.counts[pid_g8] = cnt_g8;
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
    unsigned char tval[kTableTiles];
};

// Expand one mask word: tile i of the chunk is bit 31 - i, and its ids sit at
// ids2[8 * (31 - i)], so the bit position indexes the table directly.
CB_D void expand_word( unsigned mm, const unsigned* ids2, int* rows, int& w )
{
    while ( mm )
    {
        const int p = 31 - __clz( mm );
        mm ^= 1u << p;
        rows[w++] = (int)ids2[kTileCands * p];
    }
}
CB_D void expand_word_global( unsigned mm, const unsigned* ids2, int*& out )
{
    while ( mm )
    {
        const int p = 31 - __clz( mm );
        mm ^= 1u << p;
        *out++ = (int)ids2[kTileCands * p];
    }
}

template <bool CSR>
__global__ void __launch_bounds__( kBlockT, 3 )
    k_tile_fill( const __grid_constant__ TileArgs a )
{
    extern __shared__ __align__( 16 ) unsigned char s_dyn[];
    const unsigned lane = threadIdx.x & 31u;
    const int wib = threadIdx.x >> 5;
    FillSmem& S = reinterpret_cast<FillSmem*>( s_dyn )[wib];
    const int g = (int)( lane >> 2 ), t = (int)( lane & 3u );
    const int ntiles = *a.ntiles_dev;

    for ( ;; )
    {
        unsigned tile = 0u;
        if ( lane == 0u )
            tile = atomicAdd( a.ticket + 1, 1u );
        tile = __shfl_sync( kFullMask, tile, 0 );
        if ( tile >= (unsigned)ntiles )
            break;
        const TileRec rc = unpack_rec( a.recs[tile] );
        const int chunk0 = a.chunk_off[tile];

        int pid = -1, cnt = 0;
        long long dst = 0;
        if ( (int)lane < rc.np )
        {
            pid = (int)a.permute[rc.first + lane];
            if ( pid >= a.begin && pid < a.end )
            {
                cnt = a.counts[pid];
                dst = CSR ? (long long)a.offsets[pid] : (long long)pid * a.width;
            }
        }
        uint2 sp = make_uint2( 0u, 0u );
        if ( lane < 9u )
            sp = a.spans[(size_t)tile * 9u + lane];
        // inclusive scans: the 16 row sizes; the mma tiles of the 9 spans
        int inc = cnt;
        const int my_nt = (int)( ( sp.y + kTileCands - 1 ) / kTileCands );
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
        build_tile_table( S, 0, tiles_before, my_nt, sp.x, sp.y );
        int table_seg = 0;
        __syncwarp();
        const int nchunks = ( T + kChunkTiles - 1 ) / kChunkTiles;

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
                for ( int ci = 0; ci < nchunks; ++ci )
                {
                    uint4 m = a.masks[( (size_t)chunk0 + ci ) * 32u + lane];
                    if ( !in_g )
                        m.x = m.y = 0u;
                    if ( !in_g8 )
                        m.z = m.w = 0u;
                    if ( !__any_sync( kFullMask, ( m.x | m.y | m.z | m.w ) != 0u ) )
                        continue;
                    __syncwarp();
                    const int seg = ( ci * kChunkTiles ) / kTableTiles;
                    if ( seg != table_seg )
                    {
                        build_tile_table( S, seg * kTableTiles, tiles_before, my_nt, sp.x, sp.y );
                        table_seg = seg;
                        __syncwarp();
                    }
                    // ids of the chunk's candidates.  Entries behind a span's end inside its
                    // last tile read whatever follows in `permute` (padded by 8): their
                    // mask bits are zero, the values are never used.
#pragma unroll
                    for ( int k = 0; k < kChunkEntries / 32; ++k )
                    {
                        const int e = (int)lane + 32 * k;
                        const int tl = ci * kChunkTiles + ( e >> 3 );
                        if ( tl < T )
                            S.ids[( 31 - ( e >> 3 ) ) * kTileCands + ( e & 7 )] =
                                a.permute[S.tsrc[tl & ( kTableTiles - 1 )] + (unsigned)( e & 7 )];
                    }
                    __syncwarp();
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
                        int w_g = cur_g + ( pk & 0xffff ) - pg;
                        int w_g8 = cur_g8 + ( pk >> 16 ) - pg8;
                        expand_word( m.x, ids2, S.rows, w_g );
                        expand_word( m.y, ids2 + 1, S.rows, w_g );
                        expand_word( m.z, ids2, S.rows, w_g8 );
                        expand_word( m.w, ids2 + 1, S.rows, w_g8 );
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
                    // rows leave shared memory coalesced, each to its final place
#pragma unroll 1
                    for ( int h = ha; h < hb; ++h )
                    {
                        const int c = __shfl_sync( kFullMask, cnt, h );
                        const int rs = __shfl_sync( kFullMask, rowstart, h );
                        int* d = a.neighbors + ( __shfl_sync( kFullMask, dst, h ) - rs );
                        const int end = rs + c;
#pragma unroll 1
                        for ( int i = rs + (int)lane; i < end; i += 32 )
                            __stcs( d + i, S.rows[i] );
                    }
                }
                __syncwarp();
            }
            ha = hb;
        }
    }
}

int persistent_blocks( const void* func, int smem )
{
    int nb = 0;
    if ( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &nb, func, kBlockT, smem ) !=
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
int launch_persistent( K kernel, int smem, const TileArgs& a, cudaStream_t stream )
{
    // per device/context: set unconditionally (cheap) rather than caching per process
    CB_CUDA( cudaFuncSetAttribute( kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   smem ) );
    const int blocks = persistent_blocks( (const void*)kernel, smem );
    kernel<<<blocks, kBlockT, smem, stream>>>( a );
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

double tile_filter_bound( const TileGrid& tg, double radius )
{
    // See DESIGN.md "Exactness (v2)".  u = 2^-24 (fp32 unit roundoff), v = 2^-20 (what the
    // hi/lo tf32 split drops), gam = 2^-21 per accumulated sum inside the tensor core.
    const double u = ldexp( 1.0, -24 ), v = ldexp( 1.0, -20 ), gam = ldexp( 1.0, -21 );
    double M = 0.0;
    for ( int d = 0; d < 3; ++d )
        M = fmax( M, tg.g.max[d] - tg.g.min[d] );
    M *= 1.0 + 1.0e-9;
    const double sl = 1.0 + 1.0e-5;
    const double Dx = 1.5 * tg.g.dx[0] * sl, Dy = 1.5 * tg.g.dx[1] * sl;
    const double Dz = ( 0.5 * tg.zb + tg.kz + 0.5 ) * tg.g.dx[2] * sl;
    const double Hx = 0.5 * tg.g.dx[0] * sl, Hy = 0.5 * tg.g.dx[1] * sl;
    const double Hz = ( 0.5 * tg.zb + 0.5 ) * tg.g.dx[2] * sl;
    const double Dmax = fmax( Dx, fmax( Dy, Dz ) );
    const double S = Dx * Dx + Dy * Dy + Dz * Dz;
    const double Sh = Hx * Hx + Hy * Hy + Hz * Hz;
    const double HD = Hx * Dx + Hy * Dy + Hz * Dz;
    const double rsqr = radius * radius;
    // (1)+(2) coordinate roundings (q = fl32(x - min), x' = fl32(q - O)), for s <= 4 r^2
    double E = 4.0 * sqrt( 3.0 ) * u * ( M + Dmax ) * ( 2.02 * radius ) +
               12.0 * u * u * ( M * M + Dmax * Dmax );
    // (3) norms in fp32, hi/lo splits, dropped lo*lo, cutoff split, accumulation
    E += 3.0 * u * ( S + Sh ) + v * ( S + Sh ) + 6.0 * v * HD + v * rsqr;
    E += 2.0 * u * ( Sh + rsqr ); // accumulator init fl(fl(N - r2hi) - r2lo)
    E += 2.0 * gam * ( 2.0 * HD * ( 1.0 + 1.0e-3 ) + S + Sh + rsqr );
    return E;
}

int tile_gather_q( const cb_positions& x, long long n, const unsigned* permute, float4* q,
                   const double* origin, cudaStream_t stream )
{
    k_gather_q<<<launch_grid_for( n + 8, 256 ), 256, 0, stream>>>(
        make_access( x ), n, permute, q, origin[0], origin[1], origin[2] );
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
    k_plan_tiles<<<launch_grid_for( tg.nblocks, 128 ), 128, 0, stream>>>(
        cell_off, gi, tg.zb, tg.nzb, tg.nblocks, half ? 1 : 0, block_tiles, tile_base, recs,
        spans, tile_chunks, rec_capacity );
    CB_CHECK_LAUNCH();
    CB_TRY( exclusive_scan_i32( tile_chunks, chunk_off, rec_capacity, true, nullptr,
                                scan_scratch, stream ) );
    return CB_OK;
}

static bool use_tma_staging()
{
    const char* e = getenv( "CB_TILE_STAGING" );
    return e && strcmp( e, "tma" ) == 0;
}

int tile_count_pass( const TileArgs& a, bool half, cudaStream_t stream )
{
    if ( use_tma_staging() )
    {
        const int smem = (int)sizeof( CountSmemTma ) * kWarpsT;
        return half ? launch_persistent( k_tile_count<true, false, true>, smem, a, stream )
                    : launch_persistent( k_tile_count<false, false, true>, smem, a, stream );
    }
    const int smem = (int)sizeof( CountSmemLdg ) * kWarpsT;
    return half ? launch_persistent( k_tile_count<true, false, false>, smem, a, stream )
                : launch_persistent( k_tile_count<false, false, false>, smem, a, stream );
}

int tile_diag_pass( const TileArgs& a, bool half, cudaStream_t stream )
{
    if ( use_tma_staging() )
    {
        const int smem = (int)sizeof( CountSmemTma ) * kWarpsT;
        return half ? launch_persistent( k_tile_count<true, true, true>, smem, a, stream )
                    : launch_persistent( k_tile_count<false, true, true>, smem, a, stream );
    }
    const int smem = (int)sizeof( CountSmemLdg ) * kWarpsT;
    return half ? launch_persistent( k_tile_count<true, true, false>, smem, a, stream )
                : launch_persistent( k_tile_count<false, true, false>, smem, a, stream );
}

int tile_fill_pass( const TileArgs& a, bool csr, cudaStream_t stream )
{
    const int smem = (int)sizeof( FillSmem ) * kWarpsT;
    return csr ? launch_persistent( k_tile_fill<true>, smem, a, stream )
               : launch_persistent( k_tile_fill<false>, smem, a, stream );
}

} // namespace cb
