// VerletList count / fill passes, v1: refined cell grid + FP32 filter + exact FP64 resolve.
//
// Same reference semantics as cb_verlet.cu (core/src/Cabana_VerletList.hpp:316-474 count,
// :572-713 fill) but organised for the B200's issue and L1 bandwidth limits (ncu on v0:
// ~870-1200 warp instructions per particle, issue-bound, DRAM < 6 %):
//
//  * Internal grid.  Every user cell (delta = cell_size_ratio * r) is split into m^3
//    sub-cells of ~r/2 (m a power of two), so a home cell sees ~290 candidates instead of
//    ~500.  The user cell of every particle stays EXACT (sub-cell index = user cell * m +
//    sub), so the reference's stencil range (+-R user cells,
//    Cabana_LinkedCellList.hpp:105-119) is enforced exactly, and cell-level pruning here
//    is only ever conservative (host-built table kz[|da|][|db|]).
//  * One warp owns one home cell.  Lanes first build the cell's candidate list (one lane
//    per stencil row = one contiguous span of the cell-sorted arrays, warp scan, indices
//    materialised in shared memory and padded with a far-away sentinel), then every lane
//    loads ONE candidate as a float4 and tests it against up to 4 home particles held in
//    registers: list set-up and candidate load are shared by all particles of the cell.
//  * Two-tier decision.  Tier 1 is r^2 in FP32 on origin-relative coordinates with a
//    rigorous error bound tau: s32 <= r^2 - tau is a certain hit, s32 > r^2 + tau a certain
//    miss.  Only candidates inside the band (~0.02 per particle) go to tier 2: the exact
//    reference arithmetic (__dsub_rn/__dmul_rn/__dadd_rn, fixed order, closed cutoff), the
//    exact half-list criterion, and -- when r^2 - band < s <= r^2 -- the reference's own
//    cell prune (minDistanceToPoint on the user grid, :401-403).  Every in/out decision is
//    therefore bit-identical to the reference's; DESIGN.md "Exactness" has the bounds.
//  * Hits are recorded as per-lane bit masks (no ballot/popc in the hot loop) and written
//    once per window with a warp scan, giving runs of consecutive stores per row.
#include "cb_common.cuh"
#include "cb_internal.h"
#include "cb_verlet_fine.h"

namespace cb
{
namespace
{

constexpr int kWarps = 8;
constexpr int kBlock = kWarps * 32;
constexpr int kListCap = 512; // candidates per window: 16 iterations of 32 lanes
constexpr int kCellsPerWarp = 8;
constexpr int kGroup = 4; // home particles per register group

enum
{
    kCount = 0,
    kFill = 1
};

CB_D int warp_inclusive_scan( int v, unsigned lane )
{
#pragma unroll
    for ( int o = 1; o < 32; o <<= 1 )
    {
        const int y = __shfl_up_sync( kFullMask, v, o );
        if ( (int)lane >= o )
            v += y;
    }
    return v;
}

// The reference's cell-level prune for the pair (p, n), evaluated exactly on the USER grid
// (Cabana_VerletList.hpp:401-403 -> impl/Cabana_CartesianGrid.hpp:207-223), plus the
// stencil range check (Cabana_LinkedCellList.hpp:112-118).
__device__ __noinline__ bool reference_prune_passes( const FineArgs& a, double xp,
                                                     double yp, double zp, double xn,
                                                     double yn, double zn )
{
    int cn[3];
    const double pp[3] = { xp, yp, zp };
    const double pn[3] = { xn, yn, zn };
#pragma unroll
    for ( int d = 0; d < 3; ++d )
    {
        const int cp = min( max( locate_1d( a.ug, d, pp[d] ), 0 ), a.ug.nx[d] - 1 );
        cn[d] = min( max( locate_1d( a.ug, d, pn[d] ), 0 ), a.ug.nx[d] - 1 );
        if ( abs( cn[d] - cp ) > a.R )
            return false;
    }
    return min_distance_sq( a.ug, xp, yp, zp, cn[0], cn[1], cn[2] ) <= a.rsqr;
}

// Tier 2: exact arithmetic for the lanes' ambiguous candidates of home slot `ps`.
template <bool HALF>
__device__ __noinline__ unsigned resolve_exact( const FineArgs& a, const unsigned* list,
                                                unsigned lane, unsigned ps, unsigned hit,
                                                unsigned am )
{
    const double xp = a.xs[ps];
    const double yp = a.ys[ps];
    const double zp = a.zs[ps];
    while ( am )
    {
        const int it = __ffs( am ) - 1;
        am &= am - 1;
        const unsigned idx = list[it * 32 + (int)lane];
        const double xn = a.xs[idx];
        const double yn = a.ys[idx];
        const double zn = a.zs[idx];
        bool ok = true;
        if ( HALF )
            ok = half_criterion( xp, yp, zp, xn, yn, zn );
        const double d2 = pair_dist_sq( xp, yp, zp, xn, yn, zn );
        ok = ok && ( d2 <= a.rsqr ); // closed cutoff (:254)
        if ( ok && d2 > a.rsqr - a.band )
            ok = reference_prune_passes( a, xp, yp, zp, xn, yn, zn );
        if ( !ok )
            hit &= ~( 1u << it );
    }
    return hit;
}

// Write this lane's hits of one home particle; `wr` is the lane's first slot in the row.
template <bool CSR>
__device__ __noinline__ void emit_hits( const FineArgs& a, const unsigned* list,
                                        unsigned lane, unsigned hm, long long row_base,
                                        int wr )
{
    while ( hm )
    {
        const int it = __ffs( hm ) - 1;
        hm &= hm - 1;
        const unsigned idx = list[it * 32 + (int)lane];
        const int id = __float_as_int( a.q[idx].w );
        // 2D: writes past extent(1) are dropped but still counted (:99-105)
        if ( CSR || wr < a.width )
            a.neighbors[row_base + wr] = id;
        ++wr;
    }
}

// Tier 1 hot loop: every lane tests ONE candidate per iteration against NP home particles.
// Branch-free; the list is padded to a multiple of 32 with a sentinel at infinity.
template <int NP, bool HALF>
CB_D void test_window( const float4* __restrict__ q, const unsigned* list, int nit,
                       int same, unsigned lane, float t_lo, float t_hi,
                       const float ( &xi )[kGroup], const float ( &yi )[kGroup],
                       const float ( &zi )[kGroup], unsigned ( &hit )[kGroup],
                       unsigned ( &amb )[kGroup] )
{
    unsigned bit = 1u;
    int t = (int)lane;
#pragma unroll 2
    for ( int it = 0; it < nit; ++it )
    {
        const float4 c = q[list[t]];
        const bool in_slab = HALF && ( t < same );
#pragma unroll
        for ( int p = 0; p < NP; ++p )
        {
            const float dx = xi[p] - c.x;
            const float dy = yi[p] - c.y;
            const float dz = zi[p] - c.z;
            const float s = fmaf( dz, dz, fmaf( dy, dy, dx * dx ) );
            bool in = s <= t_hi;
            bool am = s > t_lo;
            if ( HALF )
            {
                // Rows beyond the home x-slab hold only larger x (the refined cell index
                // is monotone in x): always valid.  Inside the slab float rounding is
                // monotone too: c.x > xi certainly valid, c.x < xi certainly invalid, a
                // tie goes to the exact criterion.
                in = in && ( !in_slab || c.x >= xi[p] );
                am = am || ( in_slab && c.x == xi[p] );
            }
            hit[p] |= in ? bit : 0u;
            amb[p] |= ( in && am ) ? bit : 0u;
        }
        bit <<= 1;
        t += 32;
    }
}

// General per-cell kernel.  With worklist == nullptr it walks every refined cell; otherwise
// only the cells the column kernel (below) could not take (too many stencil rows or
// candidates for its shared-memory list), read from a device-side worklist.
template <int MODE, bool HALF, bool CSR>
__global__ void __launch_bounds__( kBlock, 3 )
    k_verlet_fine( const __grid_constant__ FineArgs a, const unsigned* worklist,
                   const unsigned* work_count )
{
    __shared__ unsigned s_list[kWarps][kListCap];
    const unsigned lane = lane_id();
    const int wib = threadIdx.x >> 5;
    unsigned* list = s_list[wib];
    const long long warp = (long long)blockIdx.x * kWarps + wib;
    const long long nwarps = (long long)gridDim.x * kWarps;
    const int nxf = a.nf[0], nyf = a.nf[1], nzf = a.nf[2];
    const int lgm = a.lgm, R = a.R;
    const int mm1 = ( 1 << lgm ) - 1;
    const unsigned sentinel = (unsigned)a.n; // q[n] sits at +infinity

    const long long nwork = worklist ? (long long)*work_count : a.ncell;
    const int per_warp = worklist ? 1 : kCellsPerWarp;
    for ( long long cbase = warp * per_warp; cbase < nwork; cbase += nwarps * per_warp )
    {
        const long long first = worklist ? (long long)worklist[cbase] : cbase;
        int fc = (int)( first % nzf );
        const long long tq = first / nzf;
        int fb = (int)( tq % nyf );
        int fa = (int)( tq / nyf );
        const long long cend = worklist ? first + 1
                                        : min( first + (long long)kCellsPerWarp, a.ncell );
        for ( long long cell = first; cell < cend; ++cell )
        {
            const int ca = fa, cb_ = fb, cc = fc; // this cell's indices
            if ( ++fc == nzf )
            {
                fc = 0;
                if ( ++fb == nyf )
                {
                    fb = 0;
                    ++fa;
                }
            }
            const unsigned h0 = a.cell_off[cell];
            const unsigned h1 = a.cell_off[cell + 1];
            if ( h1 == h0 )
                continue;

            // Stencil ranges in refined cells: reference stencil (+-R user cells) cut down
            // to what the cutoff can reach (+-K refined cells), clipped to the grid.
            const int ua0 = ( ( ca >> lgm ) - R ) << lgm;
            const int ub0 = ( ( cb_ >> lgm ) - R ) << lgm;
            const int uc0 = ( ( cc >> lgm ) - R ) << lgm;
            const int ua1 = ( ( ( ca >> lgm ) + R ) << lgm ) + mm1;
            const int ub1 = ( ( ( cb_ >> lgm ) + R ) << lgm ) + mm1;
            const int uc1 = ( ( ( cc >> lgm ) + R ) << lgm ) + mm1;
            int alo = max( max( ua0, ca - a.K[0] ), 0 );
            const int ahi = min( min( ua1, ca + a.K[0] ), nxf - 1 );
            const int blo = max( max( ub0, cb_ - a.K[1] ), 0 );
            const int bhi = min( min( ub1, cb_ + a.K[1] ), nyf - 1 );
            const int zlo = max( uc0, 0 );
            const int zhi = min( uc1, nzf - 1 );
            if ( HALF )
                alo = ca; // cells with a smaller x index hold only smaller x: never valid
            const int nB = bhi - blo + 1;
            const int nrows = ( ahi - alo + 1 ) * nB;
            const int r_home = ( ca - alo ) * nB + ( cb_ - blo );
            const unsigned nB_magic = ( 65536u + (unsigned)nB - 1u ) / (unsigned)nB;

            for ( unsigned pg = h0; pg < h1; pg += kGroup )
            {
                const int np = (int)min( (unsigned)kGroup, h1 - pg );
                float xi[kGroup], yi[kGroup], zi[kGroup];
                int pid[kGroup];
                int acc[kGroup];
                unsigned active = 0u;
#pragma unroll
                for ( int p = 0; p < kGroup; ++p )
                {
                    // unused slots replicate the sentinel: they never hit
                    const float4 h = a.q[p < np ? pg + p : sentinel];
                    xi[p] = h.x;
                    yi[p] = h.y;
                    zi[p] = h.z;
                    pid[p] = __float_as_int( h.w );
                    acc[p] = 0;
                    // only rows in [begin,end) are built (:340)
                    if ( p < np && pid[p] >= a.begin && pid[p] < a.end )
                        active |= 1u << p;
                }
                if ( active == 0u )
                    continue;

                for ( int rb = 0; rb < nrows; rb += 32 )
                {
                    // lane r: one stencil row = one contiguous span of sorted slots
                    const int r = rb + (int)lane;
                    unsigned start = 0;
                    int len = 0;
                    if ( r < nrows )
                    {
                        const int ra = (int)( ( (unsigned)r * nB_magic ) >> 16 );
                        const int ap = alo + ra;
                        const int bp = blo + ( r - ra * nB );
                        const int kz = a.kz[min( abs( ap - ca ), 8 ) * 9 + min( abs( bp - cb_ ), 8 )];
                        const int zl = max( zlo, cc - kz );
                        const int zh = min( zhi, cc + kz );
                        if ( kz >= 0 && zh >= zl )
                        {
                            const int c0 = ( ap * nyf + bp ) * nzf + zl;
                            start = a.cell_off[c0];
                            len = (int)( a.cell_off[c0 + ( zh - zl ) + 1] - start );
                        }
                    }
                    const int incl = warp_inclusive_scan( len, lane );
                    const int excl = incl - len;
                    const int total = __shfl_sync( kFullMask, incl, 31 );
                    int same_total = 0;
                    if ( HALF )
                    {
                        // rows 0..nB-1 (ap == ca) come first
                        const int ns = min( max( nB - rb, 0 ), 32 );
                        if ( ns > 0 )
                            same_total = __shfl_sync( kFullMask, incl, ns - 1 );
                    }
                    // virtual list position of home particle pg (for j != i)
                    int self0 = -1;
                    {
                        const int hl = r_home - rb;
                        const unsigned hs = __shfl_sync( kFullMask, start, hl & 31 );
                        const int he = __shfl_sync( kFullMask, excl, hl & 31 );
                        if ( hl >= 0 && hl < 32 )
                            self0 = he + (int)( pg - hs );
                    }

                    for ( int w0 = 0; w0 < total; w0 += kListCap )
                    {
                        const int count = min( kListCap, total - w0 );
                        const int nit = ( count + 31 ) >> 5;
                        __syncwarp();
                        {
                            // each row lane writes its slice of the window
                            int v = max( excl, w0 );
                            const int v1 = min( incl, w0 + count );
                            unsigned val = start + (unsigned)( v - excl );
                            unsigned* dst = list + ( v - w0 );
                            int left = v1 - v;
                            while ( left >= 4 )
                            {
                                dst[0] = val;
                                dst[1] = val + 1;
                                dst[2] = val + 2;
                                dst[3] = val + 3;
                                dst += 4;
                                val += 4;
                                left -= 4;
                            }
                            if ( left > 0 )
                                dst[0] = val;
                            if ( left > 1 )
                                dst[1] = val + 1;
                            if ( left > 2 )
                                dst[2] = val + 2;
                            // pad to a multiple of 32 with the sentinel
                            if ( count + (int)lane < nit * 32 )
                                list[count + (int)lane] = sentinel;
                        }
                        __syncwarp();

                        unsigned hit[kGroup] = { 0u, 0u, 0u, 0u };
                        unsigned amb[kGroup] = { 0u, 0u, 0u, 0u };
                        const int same = HALF ? ( same_total - w0 ) : 0;
                        switch ( np )
                        {
                        case 1:
                            test_window<1, HALF>( a.q, list, nit, same, lane, a.t_lo, a.t_hi,
                                                  xi, yi, zi, hit, amb );
                            break;
                        case 2:
                            test_window<2, HALF>( a.q, list, nit, same, lane, a.t_lo, a.t_hi,
                                                  xi, yi, zi, hit, amb );
                            break;
                        case 3:
                            test_window<3, HALF>( a.q, list, nit, same, lane, a.t_lo, a.t_hi,
                                                  xi, yi, zi, hit, amb );
                            break;
                        default:
                            test_window<4, HALF>( a.q, list, nit, same, lane, a.t_lo, a.t_hi,
                                                  xi, yi, zi, hit, amb );
                            break;
                        }

                        // j != i: each home particle is in its own list exactly once
                        const int sp0 = self0 - w0;
                        if ( self0 >= 0 && sp0 + kGroup > 0 && sp0 < kListCap )
                        {
#pragma unroll
                            for ( int p = 0; p < kGroup; ++p )
                            {
                                const int sp = sp0 + p;
                                if ( sp >= 0 && sp < kListCap && ( sp & 31 ) == (int)lane )
                                {
                                    hit[p] &= ~( 1u << ( sp >> 5 ) );
                                    amb[p] &= ~( 1u << ( sp >> 5 ) );
                                }
                            }
                        }
                        // tier 2 (rare): exact arithmetic for the ambiguous band
                        if ( __any_sync( kFullMask,
                                         ( amb[0] | amb[1] | amb[2] | amb[3] ) != 0u ) )
                        {
#pragma unroll
                            for ( int p = 0; p < kGroup; ++p )
                                if ( amb[p] )
                                    hit[p] = resolve_exact<HALF>( a, list, lane, pg + p,
                                                                  hit[p], amb[p] );
                        }

                        if ( MODE == kCount )
                        {
#pragma unroll
                            for ( int p = 0; p < kGroup; ++p )
                                acc[p] += __popc( hit[p] );
                        }
                        else
                        {
#pragma unroll
                            for ( int p = 0; p < kGroup; ++p )
                            {
                                if ( p < np )
                                {
                                    const int c = __popc( hit[p] );
                                    const int inc = warp_inclusive_scan( c, lane );
                                    const int tot = __shfl_sync( kFullMask, inc, 31 );
                                    if ( ( active >> p ) & 1u )
                                    {
                                        const long long row_base =
                                            CSR ? (long long)a.offsets[pid[p]]
                                                : (long long)pid[p] * a.width;
                                        emit_hits<CSR>( a, list, lane, hit[p], row_base,
                                                        acc[p] + inc - c );
                                    }
                                    acc[p] += tot;
                                }
                            }
                        }
                    }
                }

#pragma unroll
                for ( int p = 0; p < kGroup; ++p )
                {
                    if ( ( active >> p ) & 1u )
                    {
                        const int total =
                            MODE == kCount ? warp_reduce_sum( acc[p] ) : acc[p];
                        if ( lane == 0 )
                            a.counts[pid[p]] = total;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Column kernel (the fast path).  One warp owns a run of kChunk consecutive cells of one
// (a,b) column.  The stencil rows (a',b') are the same for the whole run, so the warp builds
// ONE candidate list for it, ordered LAYER-major (z cell k, then row): the candidates of
// home cell c are then the contiguous range of layers [c-K, c+K] -- no per-cell list, no
// per-cell row geometry; each new home cell costs one extra layer (one load and one scan).
// Cells it cannot take (more than 64 stencil rows, more than 1024 candidates, or a run
// whose layers overflow the shared-memory list) go to a worklist for k_verlet_fine.
// ---------------------------------------------------------------------------------------
constexpr int kColWarps = 8;
constexpr int kColBlock = kColWarps * 32;
constexpr int kColCap = 1024;  // list entries per warp
constexpr int kChunk = 8;      // home cells per work item
constexpr int kMaxLayers = kChunk + 2 * 8 + 1;

CB_D void push_overflow( const FineArgs& a, unsigned cell, unsigned lane )
{
    if ( lane == 0 )
        a.worklist[atomicAdd( a.work_count, 1u )] = cell;
}

// Tier 1 hot loop over list positions [t0, t1).
template <int NP, bool HALF>
CB_D void test_range( const float4* __restrict__ q, const unsigned* list, int t0, int t1,
                      unsigned lane, unsigned sentinel, float t_lo, float t_hi,
                      const float ( &xi )[kGroup], const float ( &yi )[kGroup],
                      const float ( &zi )[kGroup], unsigned ( &hit )[kGroup],
                      unsigned ( &amb )[kGroup] )
{
    unsigned bit = 1u;
#pragma unroll 2
    for ( int tb = t0; tb < t1; tb += 32 )
    {
        // lanes past the end read the sentinel at infinity instead of the next layer
        const int t = tb + (int)lane;
        const unsigned idx = t < t1 ? list[t] : sentinel;
        const float4 c = q[idx];
#pragma unroll
        for ( int p = 0; p < NP; ++p )
        {
            const float dx = xi[p] - c.x;
            const float dy = yi[p] - c.y;
            const float dz = zi[p] - c.z;
            const float s = fmaf( dz, dz, fmaf( dy, dy, dx * dx ) );
            if ( HALF )
            {
                // x-major half criterion in FP32: float rounding is monotone, so c.x > xi
                // is certainly valid, c.x < xi certainly invalid, a tie goes to tier 2.
                const bool in = ( s <= t_hi ) && ( c.x >= xi[p] );
                const bool am = in && ( ( s > t_lo ) || ( c.x == xi[p] ) );
                hit[p] |= in ? bit : 0u;
                amb[p] |= am ? bit : 0u;
            }
            else
            {
                asm( "{\n\t.reg .pred p, q;\n\t"
                     "setp.le.f32 p, %2, %3;\n\t"
                     "setp.gt.and.f32 q, %2, %4, p;\n\t"
                     "@p or.b32 %0, %0, %5;\n\t"
                     "@q or.b32 %1, %1, %5;\n\t}"
                     : "+r"( hit[p] ), "+r"( amb[p] )
                     : "f"( s ), "f"( t_hi ), "f"( t_lo ), "r"( bit ) );
            }
        }
        bit <<= 1;
    }
}

template <int MODE, bool HALF, bool CSR>
__global__ void __launch_bounds__( kColBlock, 3 )
    k_verlet_column( const __grid_constant__ FineArgs a )
{
    __shared__ unsigned s_list[kColWarps][kColCap];
    __shared__ int s_layer[kColWarps][kMaxLayers + 1];
    __shared__ int s_home[kColWarps][kMaxLayers];
    const unsigned lane = threadIdx.x & 31u;
    const int wib = threadIdx.x >> 5;
    unsigned* list = s_list[wib];
    int* layer = s_layer[wib];
    int* homepos = s_home[wib];
    const int nxf = a.nf[0], nyf = a.nf[1], nzf = a.nf[2];
    const int lgm = a.lgm, R = a.R;
    const int mm1 = ( 1 << lgm ) - 1;
    const int Kz = a.K[2];
    const unsigned sentinel = (unsigned)a.n;
    const int nchunk = ( nzf + kChunk - 1 ) / kChunk;
    const long long nitems = (long long)nxf * nyf * nchunk;
    const long long warp = (long long)blockIdx.x * kColWarps + wib;
    const long long nwarps = (long long)gridDim.x * kColWarps;

    for ( long long item = warp; item < nitems; item += nwarps )
    {
        // b fastest: the warps of a block take ADJACENT columns of the same z chunk, so
        // 4/5 of their candidate columns are shared and stay hot in L1.
        const int cb_ = (int)( item % nyf );
        const long long rest = item / nyf;
        const int zc = (int)( rest % nchunk );
        const int ca = (int)( rest / nchunk );
        const int cz0 = zc * kChunk;
        const int cz1 = min( cz0 + kChunk, nzf );
        const int homebase = ( ca * nyf + cb_ ) * nzf;

        // stencil rows of this column (same arithmetic as k_verlet_fine)
        const int ua0 = ( ( ca >> lgm ) - R ) << lgm;
        const int ub0 = ( ( cb_ >> lgm ) - R ) << lgm;
        const int ua1 = ( ( ( ca >> lgm ) + R ) << lgm ) + mm1;
        const int ub1 = ( ( ( cb_ >> lgm ) + R ) << lgm ) + mm1;
        int alo = max( max( ua0, ca - a.K[0] ), 0 );
        const int ahi = min( min( ua1, ca + a.K[0] ), nxf - 1 );
        const int blo = max( max( ub0, cb_ - a.K[1] ), 0 );
        const int bhi = min( min( ub1, cb_ + a.K[1] ), nyf - 1 );
        if ( HALF )
            alo = ca;
        const int nB = bhi - blo + 1;
        const int nrows = ( ahi - alo + 1 ) * nB;
        const int r_home = ( ca - alo ) * nB + ( cb_ - blo );

        __syncwarp();
        int nbuilt = 0; // layers [kL0, kL0+nbuilt) are in the list
        const int kL0 = max( cz0 - Kz, 0 );
        const int kL1 = min( cz1 - 1 + Kz, nzf - 1 );
        if ( nrows <= 64 )
        {
            // lane owns rows `lane` and `lane + 32`
            int rowbase[2];
            bool rowok[2];
            unsigned e_prev[2];
#pragma unroll
            for ( int h = 0; h < 2; ++h )
            {
                const int r = (int)lane + 32 * h;
                rowok[h] = false;
                rowbase[h] = 0;
                e_prev[h] = 0u;
                if ( r < nrows )
                {
                    const int ra = r / nB;
                    const int ap = alo + ra;
                    const int bp = blo + ( r - ra * nB );
                    const int kz =
                        a.kz[min( abs( ap - ca ), 8 ) * 9 + min( abs( bp - cb_ ), 8 )];
                    rowok[h] = kz >= 0;
                    rowbase[h] = ( ap * nyf + bp ) * nzf;
                    if ( rowok[h] )
                        e_prev[h] = a.cell_off[rowbase[h] + kL0];
                }
            }
            int base = 0;
            for ( int k = kL0; k <= kL1; ++k )
            {
                unsigned st[2];
                int len[2];
#pragma unroll
                for ( int h = 0; h < 2; ++h )
                {
                    st[h] = e_prev[h];
                    len[h] = 0;
                    if ( rowok[h] )
                    {
                        const unsigned e = a.cell_off[rowbase[h] + k + 1];
                        len[h] = (int)( e - st[h] );
                        e_prev[h] = e;
                    }
                }
                // order inside a layer: rows 0..31 (lane order), then rows 32..63
                const int incl0 = warp_inclusive_scan( len[0], lane );
                const int tot0 = __shfl_sync( kFullMask, incl0, 31 );
                int tot = tot0;
                int excl1 = 0;
                if ( nrows > 32 )
                {
                    const int incl1 = warp_inclusive_scan( len[1], lane );
                    tot += __shfl_sync( kFullMask, incl1, 31 );
                    excl1 = tot0 + incl1 - len[1];
                }
                if ( base + tot > kColCap )
                    break;
                {
                    unsigned* dst = list + base + ( incl0 - len[0] );
                    for ( int j = 0; j < len[0]; ++j )
                        dst[j] = st[0] + (unsigned)j;
                    if ( nrows > 32 )
                    {
                        unsigned* dst1 = list + base + excl1;
                        for ( int j = 0; j < len[1]; ++j )
                            dst1[j] = st[1] + (unsigned)j;
                    }
                }
                if ( lane == 0 )
                    layer[k - kL0] = base;
                if ( (int)lane == ( r_home & 31 ) )
                    homepos[k - kL0] =
                        base + ( r_home < 32 ? incl0 - len[0] : excl1 );
                base += tot;
                ++nbuilt;
            }
            if ( lane == 0 )
                layer[nbuilt] = base;
        }
        __syncwarp();

        for ( int cc = cz0; cc < cz1; ++cc )
        {
            const unsigned h0 = a.cell_off[homebase + cc];
            const unsigned h1 = a.cell_off[homebase + cc + 1];
            if ( h1 == h0 )
                continue;
            // candidate layers of this home cell: +-Kz, inside the reference stencil
            const int uc0 = ( ( cc >> lgm ) - R ) << lgm;
            const int uc1 = ( ( ( cc >> lgm ) + R ) << lgm ) + mm1;
            const int lo = max( max( cc - Kz, uc0 ), 0 );
            const int hi = min( min( cc + Kz, uc1 ), nzf - 1 );
            if ( hi - kL0 >= nbuilt )
            {
                push_overflow( a, (unsigned)( homebase + cc ), lane );
                continue;
            }
            const int t0 = layer[lo - kL0];
            const int t1 = layer[hi + 1 - kL0];
            if ( t1 - t0 > 1024 )
            {
                push_overflow( a, (unsigned)( homebase + cc ), lane );
                continue;
            }
            const int selfbase = homepos[cc - kL0] - t0;

            for ( unsigned pg = h0; pg < h1; pg += kGroup )
            {
                const int np = (int)min( (unsigned)kGroup, h1 - pg );
                float xi[kGroup], yi[kGroup], zi[kGroup];
                int pid[kGroup];
                unsigned active = 0u;
#pragma unroll
                for ( int p = 0; p < kGroup; ++p )
                {
                    const float4 h = a.q[p < np ? pg + p : sentinel];
                    xi[p] = h.x;
                    yi[p] = h.y;
                    zi[p] = h.z;
                    pid[p] = __float_as_int( h.w );
                    // only rows in [begin,end) are built (:340); the sentinel has id -1
                    if ( pid[p] >= a.begin && pid[p] < a.end )
                        active |= 1u << p;
                }
                if ( active == 0u )
                    continue;

                unsigned hit[kGroup] = { 0u, 0u, 0u, 0u };
                unsigned amb[kGroup] = { 0u, 0u, 0u, 0u };
                switch ( np )
                {
                case 1:
                    test_range<1, HALF>( a.q, list, t0, t1, lane, sentinel, a.t_lo, a.t_hi, xi,
                                         yi, zi, hit, amb );
                    break;
                case 2:
                    test_range<2, HALF>( a.q, list, t0, t1, lane, sentinel, a.t_lo, a.t_hi, xi,
                                         yi, zi, hit, amb );
                    break;
                case 3:
                    test_range<3, HALF>( a.q, list, t0, t1, lane, sentinel, a.t_lo, a.t_hi, xi,
                                         yi, zi, hit, amb );
                    break;
                default:
                    test_range<4, HALF>( a.q, list, t0, t1, lane, sentinel, a.t_lo, a.t_hi, xi,
                                         yi, zi, hit, amb );
                    break;
                }

                // j != i: home particle pg+p sits at list position selfbase + (pg-h0) + p
                {
                    const int sp0 = selfbase + (int)( pg - h0 );
#pragma unroll
                    for ( int p = 0; p < kGroup; ++p )
                    {
                        const int sp = sp0 + p;
                        if ( ( sp & 31 ) == (int)lane )
                        {
                            hit[p] &= ~( 1u << ( sp >> 5 ) );
                            amb[p] &= ~( 1u << ( sp >> 5 ) );
                        }
                    }
                }
                // tier 2 (rare): exact arithmetic for the ambiguous band
                if ( __any_sync( kFullMask, ( amb[0] | amb[1] | amb[2] | amb[3] ) != 0u ) )
                {
#pragma unroll
                    for ( int p = 0; p < kGroup; ++p )
                        if ( amb[p] )
                            hit[p] = resolve_exact<HALF>( a, list + t0, lane, pg + p, hit[p],
                                                          amb[p] );
                }

#pragma unroll
                for ( int p = 0; p < kGroup; ++p )
                {
                    if ( p < np )
                    {
                        const int c = __popc( hit[p] );
                        if ( MODE == kCount )
                        {
                            const int total = warp_reduce_sum( c );
                            if ( lane == 0 && ( ( active >> p ) & 1u ) )
                                a.counts[pid[p]] = total;
                        }
                        else
                        {
                            const int inc = warp_inclusive_scan( c, lane );
                            const int tot = __shfl_sync( kFullMask, inc, 31 );
                            if ( ( active >> p ) & 1u )
                            {
                                const long long row_base =
                                    CSR ? (long long)a.offsets[pid[p]]
                                        : (long long)pid[p] * a.width;
                                emit_hits<CSR>( a, list + t0, lane, hit[p], row_base,
                                                inc - c );
                                if ( lane == 0 )
                                    a.counts[pid[p]] = tot;
                            }
                        }
                    }
                }
            }
        }
    }
}

template <int MODE>
int launch_mode( const FineArgs& a, int algorithm, int layout, cudaStream_t stream )
{
    if ( a.n == 0 || a.ncell == 0 )
        return CB_OK;
    const bool half = algorithm == CB_NEIGHBOR_HALF;
    const bool csr = layout == CB_LAYOUT_CSR;
    const bool columns = a.worklist != nullptr;
    if ( columns )
    {
        CB_CUDA( cudaMemsetAsync( a.work_count, 0, sizeof( unsigned ), stream ) );
        const int nchunk = ( a.nf[2] + kChunk - 1 ) / kChunk;
        long long items = (long long)a.nf[0] * a.nf[1] * nchunk;
        long long blocks = ( items + kColWarps - 1 ) / kColWarps;
        const long long cap = (long long)kNumSMs * 3 * 16;
        if ( blocks > cap )
            blocks = cap;
        const int grid = (int)blocks;
        if ( half && csr )
            k_verlet_column<MODE, true, true><<<grid, kColBlock, 0, stream>>>( a );
        else if ( half )
            k_verlet_column<MODE, true, false><<<grid, kColBlock, 0, stream>>>( a );
        else if ( csr )
            k_verlet_column<MODE, false, true><<<grid, kColBlock, 0, stream>>>( a );
        else
            k_verlet_column<MODE, false, false><<<grid, kColBlock, 0, stream>>>( a );
        CB_CHECK_LAUNCH();
    }
    // General kernel: everything (no worklist) or the column kernel's leftovers.  The
    // leftover count lives on the device; an empty worklist costs one tiny launch.
    long long blocks =
        ( a.ncell + (long long)kWarps * kCellsPerWarp - 1 ) / ( (long long)kWarps * kCellsPerWarp );
    const long long cap = columns ? (long long)kNumSMs * 3 : (long long)kNumSMs * 96;
    if ( blocks > cap )
        blocks = cap;
    const int grid = (int)blocks;
    const unsigned* wl = columns ? a.worklist : nullptr;
    const unsigned* wc = columns ? a.work_count : nullptr;
    if ( half && csr )
        k_verlet_fine<MODE, true, true><<<grid, kBlock, 0, stream>>>( a, wl, wc );
    else if ( half )
        k_verlet_fine<MODE, true, false><<<grid, kBlock, 0, stream>>>( a, wl, wc );
    else if ( csr )
        k_verlet_fine<MODE, false, true><<<grid, kBlock, 0, stream>>>( a, wl, wc );
    else
        k_verlet_fine<MODE, false, false><<<grid, kBlock, 0, stream>>>( a, wl, wc );
    CB_CHECK_LAUNCH();
    return CB_OK;
}

} // namespace

int launch_fine_pass( const FineArgs& a, bool fill, int algorithm, int layout,
                      cudaStream_t stream )
{
    return fill ? launch_mode<kFill>( a, algorithm, layout, stream )
                : launch_mode<kCount>( a, algorithm, layout, stream );
}

} // namespace cb
