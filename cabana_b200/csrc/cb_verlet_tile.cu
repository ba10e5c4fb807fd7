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

namespace cb
{
namespace
{

constexpr int kWarpsT = 8;
constexpr int kBlockT = kWarpsT * 32;
constexpr int kPieceEntries = kPieceTiles * kTileCands; // 128 candidates
constexpr int kChunkEntries = kChunkTiles * kTileCands; // 256 candidates
constexpr int kRowCap = 2048;                           // ids staged per tile in the fill pass

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
    int block, np, zlo, zhi, ntile8;
};
CB_D TileRec unpack_rec( const uint4 r )
{
    TileRec t;
    t.first = r.x;
    t.block = (int)( r.y & 0x3ffffffu );
    t.np = (int)( r.y >> 26 );
    t.zlo = (int)( r.z & 0xffffu );
    t.zhi = (int)( r.z >> 16 );
    t.ntile8 = (int)r.w;
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
                  int* __restrict__ tile_chunks, long long rec_capacity )
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
            for ( int s = 0; s < 9; ++s )
            {
                unsigned st, len;
                span_of( gi, cell_off, ca, cb_, zlo, zhi, half != 0, s, st, len );
                T += (int)( ( len + kTileCands - 1 ) / kTileCands );
            }
            if ( tb + ti < rec_capacity )
            {
                recs[tb + ti] = make_uint4(
                    first, (unsigned)b | ( ( last - first + 1u ) << 26 ),
                    (unsigned)zlo | ( (unsigned)zhi << 16 ), (unsigned)T );
                tile_chunks[tb + ti] = ( T + kChunkTiles - 1 ) / kChunkTiles;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// count pass
// ---------------------------------------------------------------------------------------
struct __align__( 16 ) CountSmem
{
    float4 raw[2][kPieceEntries]; // bulk-copy landing zone (q records), double buffered
    float4 f0[kPieceEntries];     // (hx, hy, hz, n_hi) of the staged candidates
    float4 f1[kPieceEntries];     // (lx, ly, lz, n_lo)
    unsigned sp_start[16], sp_len[16], sp_pos[16];
    unsigned long long mbar[2];
};

// Stage piece `pc` (candidate list entries [128 pc, 128 pc + 128)) into raw[buf]: padding
// is pre-filled with far-away sentinels, then every span that intersects the piece is one
// bulk copy (16-byte records, 16-byte aligned on both sides).  All lanes call.
CB_D void issue_piece( CountSmem& S, const float4* __restrict__ q, int pc, int buf,
                       unsigned lane )
{
    const float4 sent = make_float4( 1.0e18f, 1.0e18f, 1.0e18f, __int_as_float( -1 ) );
#pragma unroll
    for ( int j = 0; j < kPieceEntries / 32; ++j )
        S.raw[buf][(int)lane + 32 * j] = sent;
    fence_proxy_async(); // generic-proxy stores before the async-proxy writes
    __syncwarp();
    unsigned lo = 0u, hi = 0u, p = 0u;
    if ( lane < 9u )
    {
        p = S.sp_pos[lane];
        lo = max( p, (unsigned)( pc * kPieceEntries ) );
        hi = min( p + S.sp_len[lane], (unsigned)( ( pc + 1 ) * kPieceEntries ) );
        if ( hi < lo )
            hi = lo;
    }
    const unsigned bytes = ( hi - lo ) * 16u;
    const unsigned total = __reduce_add_sync( kFullMask, bytes );
    if ( lane == 0u )
        mbar_expect_tx( &S.mbar[buf], total );
    __syncwarp();
    if ( bytes )
        bulk_g2s( &S.raw[buf][lo - (unsigned)( pc * kPieceEntries )],
                  q + S.sp_start[lane] + ( lo - p ), bytes, &S.mbar[buf] );
}

template <bool HALF, bool DIAG>
__global__ void __launch_bounds__( kBlockT, 2 )
    k_tile_count( const __grid_constant__ TileArgs a )
{
    extern __shared__ __align__( 16 ) unsigned char s_dyn[];
    const unsigned lane = threadIdx.x & 31u;
    const int wib = threadIdx.x >> 5;
    CountSmem& S = reinterpret_cast<CountSmem*>( s_dyn )[wib];
    const int g = (int)( lane >> 2 ), t = (int)( lane & 3u );
    if ( lane == 0u )
    {
        mbar_init( &S.mbar[0], 1u );
        mbar_init( &S.mbar[1], 1u );
    }
    fence_mbar_init();
    fence_proxy_async();
    __syncwarp();
    unsigned phase = 0u; // bit b: parity the next wait on mbar[b] uses
    const int ntiles = *a.ntiles_dev;
    const GridInts gi = { a.ncx, a.ncy, a.nz, a.kz };
    const float zero4[4] = { 0.f, 0.f, 0.f, 0.f };
    float* f0w = reinterpret_cast<float*>( S.f0 );
    float* f1w = reinterpret_cast<float*>( S.f1 );

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
        const int col = rc.block / a.nzb;
        const int ca = col / a.ncy;
        const int cb_ = col - ca * a.ncy;

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
        const float Ox = ( (float)ca + 0.5f ) * a.wx;
        const float Oy = ( (float)cb_ + 0.5f ) * a.wy;
        const float Oz = ( 0.5f * (float)( rc.zlo + rc.zhi ) + 0.5f ) * a.hz;
        const float xh = hq.x - Ox, yh = hq.y - Oy, zh = hq.z - Oz;
        __syncwarp();
        if ( lane < 16u )
        {
            // A operands of the two MMAs, one row per home particle (see mma_tf32):
            //   A1 = [-2H | -2H] with 1 in slot 3 and 7;  A2 = [-2L, 0 | N_hi, N_lo, -r2hi, -r2lo]
            const float Hx = trunc_tf32( xh ), Hy = trunc_tf32( yh ), Hz = trunc_tf32( zh );
            const float N = fmaf( zh, zh, fmaf( yh, yh, xh * xh ) );
            const float Nh = trunc_tf32( N );
            S.f0[lane] = make_float4( -2.f * Hx, -2.f * Hy, -2.f * Hz, 1.f );
            S.f1[lane] = make_float4( -2.f * ( xh - Hx ), -2.f * ( yh - Hy ),
                                      -2.f * ( zh - Hz ), 0.f );
            S.raw[1][lane] = make_float4( Nh, N - Nh, -a.r2hi, -a.r2lo );
        }
        __syncwarp();
        const float a1_lo = f0w[4 * g + t], a1_hi = f0w[4 * ( g + 8 ) + t];
        const float a2_lo = f1w[4 * g + t], a2_hi = f1w[4 * ( g + 8 ) + t];
        const float a2p_lo = reinterpret_cast<float*>( S.raw[1] )[4 * g + t];
        const float a2p_hi = reinterpret_cast<float*>( S.raw[1] )[4 * ( g + 8 ) + t];
        const float hx_g = __shfl_sync( kFullMask, xh, g );
        const float hx_g8 = __shfl_sync( kFullMask, xh, g + 8 );
        const int pid_g = __shfl_sync( kFullMask, pid, g );
        const int pid_g8 = __shfl_sync( kFullMask, pid, g + 8 );
        const bool act_g = ( actmask >> g ) & 1u;
        const bool act_g8 = ( actmask >> ( g + 8 ) ) & 1u;

        // ---- candidate spans -------------------------------------------------------
        unsigned sp_start = 0u, sp_len = 0u;
        if ( lane < 9u )
            span_of( gi, a.cell_off, ca, cb_, rc.zlo, rc.zhi, HALF, (int)lane, sp_start,
                     sp_len );
        const int my_nt = (int)( ( sp_len + kTileCands - 1 ) / kTileCands );
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
        const unsigned my_pos = (unsigned)( incl - my_nt ) * kTileCands;
        __syncwarp();
        if ( lane < 16u )
        {
            S.sp_start[lane] = sp_start;
            S.sp_len[lane] = sp_len;
            S.sp_pos[lane] = my_pos;
        }
        // list position of home particle h: selfbase + h (the home span is s = 1)
        const unsigned selfbase =
            __shfl_sync( kFullMask, my_pos - sp_start, 1 ) + rc.first;
        __syncwarp();

        // ---- pieces ----------------------------------------------------------------
        const int npieces = ( T + kPieceTiles - 1 ) / kPieceTiles;
        issue_piece( S, a.q, 0, 0, lane );
        unsigned m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;
        float ma = 3.0e38f;
        int cnt_g = 0, cnt_g8 = 0;
        int in_chunk = 0, chunk_i = 0;
        for ( int pc = 0; pc < npieces; ++pc )
        {
            const int buf = pc & 1;
            if ( pc + 1 < npieces )
                issue_piece( S, a.q, pc + 1, buf ^ 1, lane );
            while ( !mbar_try_wait( &S.mbar[buf], ( phase >> buf ) & 1u ) )
            {
            }
            phase ^= 1u << buf;
            // transform: tile-local coordinates, tf32 hi/lo split, squared norm
#pragma unroll
            for ( int j = 0; j < kPieceEntries / 32; ++j )
            {
                const int e = (int)lane + 32 * j;
                const float4 r = S.raw[buf][e];
                const float x = r.x - Ox, y = r.y - Oy, z = r.z - Oz;
                const float hx = trunc_tf32( x ), hy = trunc_tf32( y ), hz = trunc_tf32( z );
                const float n = fmaf( z, z, fmaf( y, y, x * x ) );
                const float nh = trunc_tf32( n );
                S.f0[e] = make_float4( hx, hy, hz, nh );
                S.f1[e] = make_float4( x - hx, y - hy, z - hz, n - nh );
            }
            __syncwarp();
            const int nt_p = min( kPieceTiles, T - pc * kPieceTiles );
#pragma unroll 4
            for ( int j = 0; j < nt_p; ++j )
            {
                const float b0 = f0w[32 * j + (int)lane];
                const float b1 = f1w[32 * j + (int)lane];
                float c1[4], c[4];
                mma_tf32( c1, a1_lo, a1_hi, a1_lo, a1_hi, b0, b1, zero4 );
                mma_tf32( c, a2_lo, a2_hi, a2p_lo, a2p_hi, b0, 1.0f, c1 );
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
                    mma_tf32( c1, a1_lo, a1_hi, a1_lo, a1_hi, b0, b1, zero4 );
                    mma_tf32( c, a2_lo, a2_hi, a2p_lo, a2p_hi, b0, 1.0f, c1 );
                    const unsigned bit = 1u << ( nt_p - 1 - j );
#pragma unroll
                    for ( int k = 0; k < 4; ++k )
                    {
                        const int e = kTileCands * j + 2 * t + ( k & 1 );
                        const int hp = k < 2 ? pid_g : pid_g8;
                        if ( DIAG )
                        {
                            const int cp = __float_as_int( S.raw[buf][e].w );
                            if ( hp >= 0 && cp >= 0 )
                            {
                                const float err =
                                    (float)fabs( (double)c[k] - exact_c( a, hp, cp ) );
                                atomicMax( a.diag_maxerr, __float_as_uint( err ) );
                            }
                        }
                        if ( fabsf( c[k] ) <= a.tau )
                        {
                            const bool hit = exact_decide<HALF>(
                                a, hp, __float_as_int( S.raw[buf][e].w ) );
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
                            keep = exact_decide<HALF>( a, hp,
                                                       __float_as_int( S.raw[buf][e].w ) );
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
                // j != i
#pragma unroll
                for ( int hh = 0; hh < 2; ++hh )
                {
                    const unsigned p = selfbase + (unsigned)( g + 8 * hh );
                    if ( (int)( p / kChunkEntries ) == chunk_i )
                    {
                        const unsigned e = p % kChunkEntries;
                        if ( (int)( ( e & 7u ) >> 1 ) == t )
                        {
                            const unsigned bit = 0x80000000u >> ( e >> 3 );
                            if ( hh == 0 )
                            {
                                if ( e & 1u )
                                    m1 &= ~bit;
                                else
                                    m0 &= ~bit;
                            }
                            else
                            {
                                if ( e & 1u )
                                    m3 &= ~bit;
                                else
                                    m2 &= ~bit;
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
    unsigned ids[kChunkEntries];
    unsigned sp_start[16], sp_len[16], sp_pos[16];
};

template <bool HALF, bool CSR>
__global__ void __launch_bounds__( kBlockT, 2 )
    k_tile_fill( const __grid_constant__ TileArgs a )
{
    extern __shared__ __align__( 16 ) unsigned char s_dyn[];
    const unsigned lane = threadIdx.x & 31u;
    const int wib = threadIdx.x >> 5;
    FillSmem& S = reinterpret_cast<FillSmem*>( s_dyn )[wib];
    const int g = (int)( lane >> 2 ), t = (int)( lane & 3u );
    const int ntiles = *a.ntiles_dev;
    const GridInts gi = { a.ncx, a.ncy, a.nz, a.kz };

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
        const int col = rc.block / a.nzb;
        const int ca = col / a.ncy;
        const int cb_ = col - ca * a.ncy;

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
        // inclusive scan of the 16 row sizes
        int inc = cnt;
#pragma unroll
        for ( int o = 1; o < 16; o <<= 1 )
        {
            const int y = __shfl_up_sync( kFullMask, inc, o );
            if ( (int)lane >= o )
                inc += y;
        }
        const int total = __shfl_sync( kFullMask, inc, 15 );
        if ( total == 0 )
            continue;

        unsigned sp_start = 0u, sp_len = 0u;
        if ( lane < 9u )
            span_of( gi, a.cell_off, ca, cb_, rc.zlo, rc.zhi, HALF, (int)lane, sp_start,
                     sp_len );
        const int my_nt = (int)( ( sp_len + kTileCands - 1 ) / kTileCands );
        int incl = my_nt;
#pragma unroll
        for ( int o = 1; o < 16; o <<= 1 )
        {
            const int y = __shfl_up_sync( kFullMask, incl, o );
            if ( (int)lane >= o )
                incl += y;
        }
        const int T = __shfl_sync( kFullMask, incl, 15 );
        __syncwarp();
        if ( lane < 16u )
        {
            S.sp_start[lane] = sp_start;
            S.sp_len[lane] = sp_len;
            S.sp_pos[lane] = (unsigned)( incl - my_nt ) * kTileCands;
        }
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
            const int rs_g = __shfl_sync( kFullMask, rowstart, g );
            const int rs_g8 = __shfl_sync( kFullMask, rowstart, g + 8 );
            const long long dst_g = __shfl_sync( kFullMask, dst, g );
            const long long dst_g8 = __shfl_sync( kFullMask, dst, g + 8 );
            int* out_g = direct ? a.neighbors + dst_g : S.rows + rs_g;
            int* out_g8 = direct ? a.neighbors + dst_g8 : S.rows + rs_g8;
            int cur_g = 0, cur_g8 = 0;
            const int wtotal = __shfl_sync( kFullMask, inc, hb - 1 ) - base;

            if ( wtotal > 0 )
            {
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
                    // ids of the chunk's candidates
                    for ( int s = 0; s < 9; ++s )
                    {
                        const unsigned p = S.sp_pos[s];
                        const unsigned lo = max( p, (unsigned)( ci * kChunkEntries ) );
                        const unsigned hi =
                            min( p + S.sp_len[s], (unsigned)( ( ci + 1 ) * kChunkEntries ) );
                        const unsigned src = S.sp_start[s] - p;
                        for ( unsigned e = lo + lane; e < hi; e += 32u )
                            S.ids[e - (unsigned)( ci * kChunkEntries )] = a.permute[src + e];
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
                    int w_g = cur_g + ( pk & 0xffff ) - pg;
                    int w_g8 = cur_g8 + ( pk >> 16 ) - pg8;
                    cur_g += qt & 0xffff;
                    cur_g8 += qt >> 16;
                    // tile i of the chunk is bit 31 - i; candidate 2t (+1) of that tile
                    unsigned mm = m.x;
                    while ( mm )
                    {
                        const int i = __clz( mm );
                        mm &= ~( 0x80000000u >> i );
                        out_g[w_g++] = (int)S.ids[kTileCands * i + 2 * t];
                    }
                    mm = m.y;
                    while ( mm )
                    {
                        const int i = __clz( mm );
                        mm &= ~( 0x80000000u >> i );
                        out_g[w_g++] = (int)S.ids[kTileCands * i + 2 * t + 1];
                    }
                    mm = m.z;
                    while ( mm )
                    {
                        const int i = __clz( mm );
                        mm &= ~( 0x80000000u >> i );
                        out_g8[w_g8++] = (int)S.ids[kTileCands * i + 2 * t];
                    }
                    mm = m.w;
                    while ( mm )
                    {
                        const int i = __clz( mm );
                        mm &= ~( 0x80000000u >> i );
                        out_g8[w_g8++] = (int)S.ids[kTileCands * i + 2 * t + 1];
                    }
                }
                __syncwarp();
                if ( !direct )
                {
                    // rows leave shared memory coalesced, each to its final place
                    for ( int h = ha; h < hb; ++h )
                    {
                        const int c = __shfl_sync( kFullMask, cnt, h );
                        const int rs = __shfl_sync( kFullMask, rowstart, h );
                        const long long d = __shfl_sync( kFullMask, dst, h );
                        for ( int i = (int)lane; i < c; i += 32 )
                            __stcs( a.neighbors + d + i, S.rows[rs + i] );
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
    // block ids share a word with the tile's particle count (26 bits)
    while ( tg.ncols * tg.nzb >= ( 1ll << 26 ) )
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
               int* tile_base, uint4* recs, int* tile_chunks, int* chunk_off,
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
        tile_chunks, rec_capacity );
    CB_CHECK_LAUNCH();
    CB_TRY( exclusive_scan_i32( tile_chunks, chunk_off, rec_capacity, true, nullptr,
                                scan_scratch, stream ) );
    return CB_OK;
}

int tile_count_pass( const TileArgs& a, bool half, cudaStream_t stream )
{
    const int smem = (int)sizeof( CountSmem ) * kWarpsT;
    return half ? launch_persistent( k_tile_count<true, false>, smem, a, stream )
                : launch_persistent( k_tile_count<false, false>, smem, a, stream );
}

int tile_diag_pass( const TileArgs& a, bool half, cudaStream_t stream )
{
    const int smem = (int)sizeof( CountSmem ) * kWarpsT;
    return half ? launch_persistent( k_tile_count<true, true>, smem, a, stream )
                : launch_persistent( k_tile_count<false, true>, smem, a, stream );
}

int tile_fill_pass( const TileArgs& a, bool half, bool csr, cudaStream_t stream )
{
    const int smem = (int)sizeof( FillSmem ) * kWarpsT;
    if ( half )
        return csr ? launch_persistent( k_tile_fill<true, true>, smem, a, stream )
                   : launch_persistent( k_tile_fill<true, false>, smem, a, stream );
    return csr ? launch_persistent( k_tile_fill<false, true>, smem, a, stream )
               : launch_persistent( k_tile_fill<false, false>, smem, a, stream );
}

} // namespace cb
