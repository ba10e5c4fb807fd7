// VerletList construction for sm_100a.
//
// Replaces the device work of VerletList::build (core/src/Cabana_VerletList.hpp):
//   VerletListBuilder::init                    :205-240  (LCL over ALL particles)
//   count functor + linkedcell_stencil_count   :316-474  (HOT LOOP A)
//   processCounts(CSR) / processCounts(2D)     :507-562
//   fill functor + linkedcell_stencil_fill     :572-713  (HOT LOOP B)
//   buildImpl orchestration                    :1441-1484
//
// Design (v0).  Particles are binned on the reference's own grid (delta =
// cell_size_ratio * radius) and their positions gathered once into cell-sorted SoA
// arrays, so a run of stencil cells along z -- cells are numbered z-fastest -- is ONE
// contiguous span of candidates.  One warp owns one particle: lanes first evaluate the
// reference's cell-level prune (minDistanceToPoint <= r^2, exact arithmetic) for the
// (2R+1)^2 stencil rows in parallel, then sweep each surviving row span 32 candidates
// at a time with coalesced loads, test r^2 in the reference's operation order with
// __dsub_rn/__dmul_rn/__dadd_rn (never FMA), and compact hits with ballot/popc so a
// row's ids are written with coalesced stores.  The candidate set is exactly the
// reference's, so every in/out decision matches bit for bit by construction.
#include <new>
#include <utility>

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "cb_common.cuh"
#include "cb_internal.h"
#include "cb_verlet_fine.h"
#include "cb_verlet_tile.h"

namespace cb
{

namespace
{

constexpr int kBlock = 256;

struct VerletArgs
{
    const double* xs;
    const double* ys;
    const double* zs;
    const unsigned* ids;      // sorted slot -> particle id (the LCL permutation)
    const unsigned* cell_off; // [ncell+1]
    Grid g;                   // binning grid == stencil grid (Appendix A.1)
    int cell_range;
    double rsqr;
    long long n;
    long long begin, end;
    int* counts;
    const int* offsets; // CSR row starts (fill pass)
    int* neighbors;
    long long width; // 2D allocated row width
};

__global__ void __launch_bounds__( kBlock )
    k_gather_sorted( PosAccess x, long long n, const unsigned* __restrict__ permute,
                     double* __restrict__ xs, double* __restrict__ ys,
                     double* __restrict__ zs, float4* __restrict__ q, double ox, double oy,
                     double oz )
{
    for ( long long s = (long long)blockIdx.x * kBlock + threadIdx.x; s < n;
          s += (long long)gridDim.x * kBlock )
    {
        const unsigned pid = permute[s];
        const long long off = x.offset( (long long)pid );
        const double px = x.base[off];
        const double py = x.base[off + x.comp_stride];
        const double pz = x.base[off + 2 * x.comp_stride];
        xs[s] = px;
        ys[s] = py;
        zs[s] = pz;
        if ( q )
        {
            // origin-relative FP32 copy for the tier-1 filter; w carries the particle id
            q[s] = make_float4( __double2float_rn( px - ox ), __double2float_rn( py - oy ),
                                __double2float_rn( pz - oz ), __int_as_float( (int)pid ) );
            if ( s == 0 ) // sentinel used to pad candidate lists: never within any cutoff
                q[n] = make_float4( 3.0e38f, 3.0e38f, 3.0e38f, __int_as_float( -1 ) );
        }
    }
}

enum
{
    kCount = 0,
    kFill = 1
};

template <int MODE, bool HALF, bool CSR>
__global__ void __launch_bounds__( kBlock ) k_verlet_pass( const VerletArgs a )
{
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    const long long warp =
        ( (long long)blockIdx.x * kBlock + threadIdx.x ) >> 5;
    const long long nwarps = ( (long long)gridDim.x * kBlock ) >> 5;
    const int R = a.cell_range;

    for ( long long s = warp; s < a.n; s += nwarps )
    {
        const long long pid = (long long)a.ids[s];
        // Only rows in [begin,end) are built; everyone is a candidate (:340).
        if ( pid < a.begin || pid >= a.end )
            continue;
        const double xp = a.xs[s];
        const double yp = a.ys[s];
        const double zp = a.zs[s];

        // Home cell: same arithmetic as the binning kernel.
        int ci = locate_1d( a.g, 0, xp );
        int cj = locate_1d( a.g, 1, yp );
        int ck = locate_1d( a.g, 2, zp );
        ci = min( max( ci, 0 ), a.g.nx[0] - 1 );
        cj = min( max( cj, 0 ), a.g.nx[1] - 1 );
        ck = min( max( ck, 0 ), a.g.nx[2] - 1 );

        // LinkedCellStencil::getCells (Cabana_LinkedCellList.hpp:105-119).  For half
        // lists rows with i < ci are skipped: locatePoint is monotone in x, so every
        // candidate there has xn < xp and fails isValid (Cabana_NeighborList.hpp:144).
        const int imin = HALF ? ci : max( ci - R, 0 );
        const int imax = min( ci + R + 1, a.g.nx[0] );
        const int jmin = max( cj - R, 0 );
        const int jmax = min( cj + R + 1, a.g.nx[1] );
        const int kmin = max( ck - R, 0 );
        const int kmax = min( ck + R + 1, a.g.nx[2] );
        const int nj = jmax - jmin;
        const int nrow = ( imax - imin ) * nj;

        long long row_base = 0;
        if ( MODE == kFill )
            row_base = CSR ? (long long)a.offsets[pid] : pid * a.width;
        int total = 0;

        for ( int rb = 0; rb < nrow; rb += 32 )
        {
            // Lane r evaluates the cell prune for stencil row r (:401-403/:648-650).
            const int r = rb + (int)lane;
            unsigned st = 0, en = 0;
            if ( r < nrow )
            {
                const int ia = imin + r / nj;
                const int jb = jmin + r % nj;
                const double sxy = CB_ADD( min_dist_1d_sq( a.g, 0, xp, ia ),
                                           min_dist_1d_sq( a.g, 1, yp, jb ) );
                int k0 = -1, k1 = -1;
                for ( int k = kmin; k < kmax; ++k )
                {
                    const double sq = CB_ADD( sxy, min_dist_1d_sq( a.g, 2, zp, k ) );
                    if ( sq <= a.rsqr )
                    {
                        if ( k0 < 0 )
                            k0 = k;
                        k1 = k;
                    }
                }
                if ( k0 >= 0 )
                {
                    const int c0 = cardinal_index( a.g, ia, jb, k0 );
                    st = a.cell_off[c0];
                    en = a.cell_off[c0 + ( k1 - k0 ) + 1];
                }
            }
            unsigned live = __ballot_sync( kFullMask, en > st );
            while ( live )
            {
                const int rr = __ffs( live ) - 1;
                live &= live - 1;
                const unsigned rst = __shfl_sync( kFullMask, st, rr );
                const unsigned ren = __shfl_sync( kFullMask, en, rr );
                for ( unsigned j0 = rst; j0 < ren; j0 += 32 )
                {
                    const unsigned j = j0 + lane;
                    bool hit = false;
                    if ( j < ren )
                    {
                        const double xn = a.xs[j];
                        const double yn = a.ys[j];
                        const double zn = a.zs[j];
                        // p != n  <=>  different sorted slot.
                        bool ok = ( (long long)j != s );
                        if ( HALF )
                            ok = ok && half_criterion( xp, yp, zp, xn, yn, zn );
                        const double d2 = pair_dist_sq( xp, yp, zp, xn, yn, zn );
                        hit = ok && ( d2 <= a.rsqr ); // withinCutoff :254
                    }
                    const unsigned m = __ballot_sync( kFullMask, hit );
                    if ( MODE == kFill )
                    {
                        if ( hit )
                        {
                            const int pos = total + __popc( m & lt );
                            // 2D: writes past extent(1) are dropped but still
                            // counted (:99-105).
                            if ( CSR || pos < a.width )
                                a.neighbors[row_base + pos] = (int)a.ids[j];
                        }
                    }
                    total += __popc( m );
                }
            }
        }
        if ( lane == 0 )
            a.counts[pid] = total;
    }
}

// ---------------------------------------------------------------------------------------
// Per-particle cutoff radius (Cabana_VerletList.hpp:181-203, :244-305).  The linked cells and
// the stencil come from the background radius; the cell prune (:401-403) and the pair test
// (:254) use radius(i) of the particle whose row is built; a pair with dist_sqr >=
// radius(j)^2 will not be found from j's side, so i is appended to j's row as well (:299-304).
// Rows are the reference's FILL semantics and counts/offsets describe exactly those rows
// (the reference's count pass books the extra entry on the wrong row, SURVEY.md Appendix B.6;
// its own test checks the post-fill counts: 6 / 4 in tstNeighborList.hpp:244-252).
// Reference-shaped kernel (warp per particle, exact FP64, count then fill): this is a
// widening row, not the hot path of the benchmark.
// ---------------------------------------------------------------------------------------
struct RadiiArgs
{
    VerletArgs v;
    FieldAccess radii;
    int* cursor; // fill position of every row (atomic)
};

CB_D double radius_sq( const FieldAccess& r, long long i )
{
    const double v = reinterpret_cast<const double*>( r.base )[r.offset( i )];
    return CB_MUL( v, v );
}

template <int MODE, bool HALF, bool CSR>
__global__ void __launch_bounds__( kBlock ) k_verlet_radii_pass( const RadiiArgs ra )
{
    const VerletArgs& a = ra.v;
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    const long long warp = ( (long long)blockIdx.x * kBlock + threadIdx.x ) >> 5;
    const long long nwarps = ( (long long)gridDim.x * kBlock ) >> 5;
    const int R = a.cell_range;

    for ( long long s = warp; s < a.n; s += nwarps )
    {
        const long long pid = (long long)a.ids[s];
        if ( pid < a.begin || pid >= a.end )
            continue;
        const double xp = a.xs[s], yp = a.ys[s], zp = a.zs[s];
        const double ri2 = radius_sq( ra.radii, pid );
        int ci = locate_1d( a.g, 0, xp );
        int cj = locate_1d( a.g, 1, yp );
        int ck = locate_1d( a.g, 2, zp );
        ci = min( max( ci, 0 ), a.g.nx[0] - 1 );
        cj = min( max( cj, 0 ), a.g.nx[1] - 1 );
        ck = min( max( ck, 0 ), a.g.nx[2] - 1 );
        const int imin = HALF ? ci : max( ci - R, 0 );
        const int imax = min( ci + R + 1, a.g.nx[0] );
        const int jmin = max( cj - R, 0 );
        const int jmax = min( cj + R + 1, a.g.nx[1] );
        const int kmin = max( ck - R, 0 );
        const int kmax = min( ck + R + 1, a.g.nx[2] );
        const int nj = jmax - jmin;
        const int nrow = ( imax - imin ) * nj;
        const long long row_base = CSR ? ( MODE == kFill ? (long long)a.offsets[pid] : 0 )
                                       : pid * a.width;
        int total = 0;

        for ( int rb = 0; rb < nrow; rb += 32 )
        {
            const int r = rb + (int)lane;
            unsigned st = 0, en = 0;
            if ( r < nrow )
            {
                const int ia = imin + r / nj;
                const int jb = jmin + r % nj;
                const double sxy = CB_ADD( min_dist_1d_sq( a.g, 0, xp, ia ),
                                           min_dist_1d_sq( a.g, 1, yp, jb ) );
                int k0 = -1, k1 = -1;
                for ( int k = kmin; k < kmax; ++k )
                {
                    const double sq = CB_ADD( sxy, min_dist_1d_sq( a.g, 2, zp, k ) );
                    if ( sq <= ri2 ) // withinCutoff( pid, minDistanceToPoint ) :401-403
                    {
                        if ( k0 < 0 )
                            k0 = k;
                        k1 = k;
                    }
                }
                if ( k0 >= 0 )
                {
                    const int c0 = cardinal_index( a.g, ia, jb, k0 );
                    st = a.cell_off[c0];
                    en = a.cell_off[c0 + ( k1 - k0 ) + 1];
                }
            }
            unsigned live = __ballot_sync( kFullMask, en > st );
            while ( live )
            {
                const int rr = __ffs( live ) - 1;
                live &= live - 1;
                const unsigned rst = __shfl_sync( kFullMask, st, rr );
                const unsigned ren = __shfl_sync( kFullMask, en, rr );
                for ( unsigned j0 = rst; j0 < ren; j0 += 32 )
                {
                    const unsigned j = j0 + lane;
                    bool hit = false, extra = false;
                    int nid = -1;
                    if ( j < ren )
                    {
                        const double xn = a.xs[j], yn = a.ys[j], zn = a.zs[j];
                        nid = (int)a.ids[j];
                        bool ok = ( (long long)j != s );
                        if ( HALF )
                            ok = ok && half_criterion( xp, yp, zp, xn, yn, zn );
                        const double d2 = pair_dist_sq( xp, yp, zp, xn, yn, zn );
                        hit = ok && ( d2 <= ri2 );                                 // :299
                        extra = hit && ( d2 >= radius_sq( ra.radii, nid ) );       // :268
                    }
                    const unsigned m = __ballot_sync( kFullMask, hit );
                    if ( MODE == kCount )
                    {
                        if ( extra )
                            atomicAdd( a.counts + nid, 1 ); // the row the fill appends to (:304)
                    }
                    else
                    {
                        int base = 0;
                        if ( m != 0u && lane == (unsigned)( __ffs( m ) - 1 ) )
                            base = atomicAdd( ra.cursor + pid, __popc( m ) );
                        base = __shfl_sync( kFullMask, base, m ? __ffs( m ) - 1 : 0 );
                        if ( hit )
                        {
                            const int pos = base + __popc( m & lt );
                            if ( CSR || pos < a.width )
                                a.neighbors[row_base + pos] = nid;
                        }
                        if ( extra )
                        {
                            const int pos = atomicAdd( ra.cursor + nid, 1 );
                            const long long nb = CSR ? (long long)a.offsets[nid]
                                                     : (long long)nid * a.width;
                            if ( CSR || pos < a.width )
                                a.neighbors[nb + pos] = (int)pid;
                        }
                    }
                    total += __popc( m );
                }
            }
        }
        if ( MODE == kCount && lane == 0 && total > 0 )
            atomicAdd( a.counts + pid, total );
    }
}

template <int MODE>
int launch_radii_pass( const RadiiArgs& a, int algorithm, int layout, cudaStream_t stream )
{
    if ( a.v.n == 0 )
        return CB_OK;
    long long blocks = ( a.v.n * 32 + kBlock - 1 ) / kBlock;
    const long long cap = (long long)kNumSMs * 64;
    if ( blocks > cap )
        blocks = cap;
    const int grid = (int)blocks;
    const bool half = algorithm == CB_NEIGHBOR_HALF;
    const bool csr = layout == CB_LAYOUT_CSR;
    if ( half && csr )
        k_verlet_radii_pass<MODE, true, true><<<grid, kBlock, 0, stream>>>( a );
    else if ( half )
        k_verlet_radii_pass<MODE, true, false><<<grid, kBlock, 0, stream>>>( a );
    else if ( csr )
        k_verlet_radii_pass<MODE, false, true><<<grid, kBlock, 0, stream>>>( a );
    else
        k_verlet_radii_pass<MODE, false, false><<<grid, kBlock, 0, stream>>>( a );
    CB_CHECK_LAUNCH();
    return CB_OK;
}

// NumSpaceDim = 2 (Cabana_VerletList.hpp:377-392, :626-639): the two coordinates are copied
// into a dense (n,3) view with z = 0.
__global__ void __launch_bounds__( kBlock )
    k_embed_2d( PosAccess x, long long n, double* __restrict__ out )
{
    for ( long long p = (long long)blockIdx.x * kBlock + threadIdx.x; p < n;
          p += (long long)gridDim.x * kBlock )
    {
        const long long off = x.offset( p );
        out[3 * p] = x.base[off];
        out[3 * p + 1] = x.base[off + x.comp_stride];
        out[3 * p + 2] = 0.0;
    }
}

__global__ void k_set_neighbor( int* neighbors, const int* offsets, long long i,
                                long long k, long long width, int value )
{
    const long long at = offsets ? (long long)offsets[i] + k : i * width + k;
    neighbors[at] = value;
}

template <int MODE>
int launch_pass( const VerletArgs& a, int algorithm, int layout, cudaStream_t stream )
{
    if ( a.n == 0 )
        return CB_OK;
    // one warp per particle; grid capped to a multiple of the SM count
    long long blocks = ( a.n * 32 + kBlock - 1 ) / kBlock;
    const long long cap = (long long)kNumSMs * 64;
    if ( blocks > cap )
        blocks = cap;
    const int grid = (int)blocks;
    const bool half = algorithm == CB_NEIGHBOR_HALF;
    const bool csr = layout == CB_LAYOUT_CSR;
    if ( half && csr )
        k_verlet_pass<MODE, true, true><<<grid, kBlock, 0, stream>>>( a );
    else if ( half )
        k_verlet_pass<MODE, true, false><<<grid, kBlock, 0, stream>>>( a );
    else if ( csr )
        k_verlet_pass<MODE, false, true><<<grid, kBlock, 0, stream>>>( a );
    else
        k_verlet_pass<MODE, false, false><<<grid, kBlock, 0, stream>>>( a );
    CB_CHECK_LAUNCH();
    return CB_OK;
}

} // namespace
} // namespace cb

// =====================================================================================
// C ABI
// =====================================================================================
using namespace cb;

struct cb_verlet
{
    // results
    int layout = CB_LAYOUT_CSR;
    int algorithm = CB_NEIGHBOR_FULL;
    long long n = 0;
    long long total = 0;
    long long max_n = 0;
    long long width = 0;
    long long extent = 0;
    int refilled = 0;
    int row_placement = CB_ROWS_REFERENCE;
    bool built = false;
    DeviceBuffer counts, offsets, neighbors;
    // workspace
    DeviceBuffer cell_counts, cell_off, permute, cell_of, rank, scan, xs, ys, zs, q, stats;
    DeviceBuffer worklist, tmp, tmp_off, ctrl;
    // v2 (tile) workspace
    DeviceBuffer block_tiles, tile_base, recs, spans, tile_chunks, chunk_off, masks, cellslot,
        pads, cnt_sorted, dst_sorted;
    DeviceBuffer row_cursor; // per-row fill positions (per-particle radius build)
    DeviceBuffer host_stage; // device copy of host positions (build_host)
    PinnedScalars pinned;
    cudaEvent_t ev_stats = nullptr; // "sizes are on the host" (the build's one host wait)
    // optional phase timing: a ring of event sets, one per build, so a caller can time many
    // builds back to back and read the averages once at the end (no host wait per build)
    static constexpr int kProfRing = 32;
    bool profiling = false;
    cudaEvent_t ev[kProfRing][CB_VERLET_NUM_PHASES + 1] = {};
    bool ev_valid[kProfRing][CB_VERLET_NUM_PHASES + 1] = {};
    bool have_events = false;
    int prof_slot = 0;
    ~cb_verlet()
    {
        if ( have_events )
            for ( auto& set : ev )
                for ( auto& e : set )
                    cudaEventDestroy( e );
        if ( ev_stats )
            cudaEventDestroy( ev_stats );
    }
    void reset_profile()
    {
        for ( auto& set : ev_valid )
            for ( auto& f : set )
                f = false;
        prof_slot = 0;
    }
    // a build starts: take the next event set of the ring
    void begin_profile()
    {
        prof_slot = ( prof_slot + 1 ) % kProfRing;
        for ( auto& f : ev_valid[prof_slot] )
            f = false;
    }
    // mark(i): boundary i on the stream -- 0 start, 1 after binning, 2 after gather,
    // 3 after count, 4 after scan, 5 after fill.
    void mark( int i, cudaStream_t s )
    {
        if ( profiling && have_events )
        {
            cudaEventRecord( ev[prof_slot][i], s );
            ev_valid[prof_slot][i] = true;
        }
    }
};

extern "C" int cb_verlet_create( cb_verlet** out )
{
    if ( !out )
        return fail( CB_ERR_INVALID, "cb_verlet_create: null argument" );
    cb_verlet* v = new ( std::nothrow ) cb_verlet();
    if ( !v )
        return fail( CB_ERR_NOMEM, "cb_verlet_create" );
    *out = v;
    return CB_OK;
}

extern "C" int cb_verlet_destroy( cb_verlet* v )
{
    delete v;
    return CB_OK;
}


// ---------------------------------------------------------------------------------------
// v2 build: internal pencil grid -> bin -> gather q -> plan tiles -> count pass (tensor-core
// distance tiles, hit masks) -> processCounts (:507-562) -> fill pass.  diag != nullptr runs
// the filter self-test instead of the count pass: diag[0] = max |c_mma - c_exact| seen,
// diag[1] = the bound tau/2 the decisions rely on.
// ---------------------------------------------------------------------------------------
static int build_tile( cb_verlet* v, const cb_positions* x, long long begin, long long end,
                       double radius, const cb_grid& ugrid, int cell_range,
                       const double* grid_min, const double* grid_max, long long max_neigh,
                       int algorithm, int layout, cudaStream_t stream, double* diag,
                       const long long* n_dev = nullptr )
{
    // n_dev (device, optional): the particle count when the host only knows the bound x->n
    // (cb_slab_step: ghost counts stay on the GPU).  Every array is sized for x->n; rows and
    // slots past the real count stay empty; v->n is set from the read-back below.
    const long long n = x->n;
    const size_t na = (size_t)( n > 0 ? n : 1 );
    const bool half = algorithm == CB_NEIGHBOR_HALF;
    const bool csr = layout == CB_LAYOUT_CSR;
    const double rsqr = radius * radius; // :239

    TileGrid tg;
    make_tile_grid( tg, grid_min, grid_max, radius, n );
    if ( getenv( "CB_TILE_DEBUG" ) )
        fprintf( stderr,
                 "[cb] pencil grid %d x %d x %d, dx %.6f %.6f %.6f, x in [%.6f, %.6f], kz %d\n",
                 tg.ncx, tg.ncy, tg.nz, tg.g.dx[0], tg.g.dx[1], tg.g.dx[2], tg.g.min[0],
                 tg.g.max[0], tg.kz );
    const long long ns_cap = sorted_capacity( tg, n ); // sorted slots incl. column pads
    const long long rec_capacity = ns_cap / kTileHomes + tg.nblocks + 2;

    CB_TRY( v->pinned.ensure() );
    CB_TRY( v->stats.ensure( 4 * sizeof( long long ) ) );
    CB_TRY( v->ctrl.ensure( 64 ) );
    CB_TRY( v->counts.ensure( sizeof( int ) * na, 1.1 ) );
    if ( csr )
        CB_TRY( v->offsets.ensure( sizeof( int ) * ( na + 1 ), 1.1 ) );
    CB_TRY( v->cell_counts.ensure( sizeof( int ) * (size_t)tg.ncells ) );
    CB_TRY( v->cell_off.ensure( sizeof( unsigned ) * (size_t)( tg.ncells + 1 ) ) );
    CB_TRY( v->permute.ensure( sizeof( unsigned ) * (size_t)( ns_cap + 8 ), 1.1 ) );
    CB_TRY( v->q.ensure( sizeof( float4 ) * (size_t)( ns_cap + 8 ), 1.1 ) );
    CB_TRY( v->cellslot.ensure( sizeof( uint2 ) * na, 1.1 ) );
    CB_TRY( v->cnt_sorted.ensure( sizeof( int ) * (size_t)( ns_cap + 8 ), 1.1 ) );
    if ( csr )
        CB_TRY( v->dst_sorted.ensure( sizeof( int ) * (size_t)( ns_cap + 8 ), 1.1 ) );
    CB_TRY( v->pads.ensure( (size_t)tg.ncols + 16 ) );
    CB_TRY( v->block_tiles.ensure( sizeof( int ) * (size_t)( tg.nblocks + 1 ) ) );
    CB_TRY( v->tile_base.ensure( sizeof( int ) * (size_t)( tg.nblocks + 1 ) ) );
    CB_TRY( v->recs.ensure( sizeof( uint4 ) * (size_t)rec_capacity, 1.1 ) );
    CB_TRY( v->spans.ensure( sizeof( uint2 ) * 9 * (size_t)rec_capacity, 1.1 ) );
    CB_TRY( v->tile_chunks.ensure( sizeof( int ) * (size_t)rec_capacity, 1.1 ) );
    CB_TRY( v->chunk_off.ensure( sizeof( int ) * (size_t)( rec_capacity + 1 ), 1.1 ) );

    v->begin_profile();
    v->mark( 0, stream );
    CB_CUDA( cudaMemsetAsync( v->counts.ptr, 0, sizeof( int ) * na, stream ) ); // :215-216

    // LCL over ALL particles, not just [begin,end) (:229-235), on the internal grid
    CB_TRY( tile_bin( tg, *x, v->cell_counts.as<int>(), v->cell_off.as<unsigned>(),
                      v->cellslot.as<uint2>(), v->pads.as<unsigned char>(), v->q.as<float4>(),
                      v->permute.as<unsigned>(), v->scan, stream, n_dev ) );
    v->mark( 1, stream );
    CB_TRY( tile_plan( tg, v->cell_off.as<unsigned>(), half, v->block_tiles.as<int>(),
                       v->tile_base.as<int>(), v->recs.as<uint4>(), v->spans.as<uint2>(),
                       v->tile_chunks.as<int>(), v->chunk_off.as<int>(), rec_capacity,
                       v->scan, stream ) );
    v->mark( 2, stream );

    TileArgs a;
    memset( &a, 0, sizeof( a ) );
    a.q = v->q.as<float4>();
    a.permute = v->permute.as<unsigned>();
    a.cell_off = v->cell_off.as<unsigned>();
    a.x = make_access( *x );
    a.ncx = tg.ncx;
    a.ncy = tg.ncy;
    a.nz = tg.nz;
    a.zb = tg.zb;
    a.nzb = tg.nzb;
    a.kz = tg.kz;
    a.wx = (float)tg.g.dx[0];
    a.wy = (float)tg.g.dx[1];
    a.hz = (float)tg.g.dx[2];
    a.recs = v->recs.as<uint4>();
    a.spans = v->spans.as<uint2>();
    a.chunk_off = v->chunk_off.as<int>();
    a.ntiles_dev = v->tile_base.as<int>() + tg.nblocks;
    a.ticket = reinterpret_cast<unsigned*>( v->ctrl.as<char>() );
    a.overflow = reinterpret_cast<int*>( v->ctrl.as<char>() + 8 );
    a.diag_maxerr = reinterpret_cast<unsigned*>( v->ctrl.as<char>() + 16 );
    a.ug = to_grid( ugrid );
    a.R = cell_range;
    a.rsqr = rsqr;
    {
        double extent = 0.0, cmax = 0.0;
        for ( int d = 0; d < 3; ++d )
        {
            extent = fmax( extent, grid_max[d] - grid_min[d] );
            cmax = fmax( cmax, fmax( fabs( grid_min[d] ), fabs( grid_max[d] ) ) );
        }
        // Band in which the reference's cell prune can reject an in-range pair: its min
        // distance is computed with O(ulp(coordinate)) error.
        const double eta = ldexp( cmax + extent + radius, -46 );
        a.band = 8.0 * radius * eta + 4.0 * eta * eta + ldexp( rsqr, -44 );
        // tf32 split of r^2 (both parts exactly representable)
        float hi = (float)rsqr;
        unsigned hb;
        memcpy( &hb, &hi, 4 );
        hb &= 0xffffe000u;
        memcpy( &hi, &hb, 4 );
        float lo = (float)( rsqr - (double)hi );
        memcpy( &hb, &lo, 4 );
        hb &= 0xffffe000u;
        memcpy( &lo, &hb, 4 );
        a.r2hi = hi;
        a.r2lo = lo;
        // tau = twice the proven bound for a tile whose home particles span up to zb z cells
        a.tau = nextafterf( (float)( 2.0 * tile_filter_bound( tg, radius, tg.zb ) ), INFINITY );
        if ( diag )
            diag[1] = tile_filter_bound( tg, radius, tg.zb );
    }
    a.n = n;
    a.begin = begin;
    a.end = end;
    a.counts = v->counts.as<int>();
    a.cnt_sorted = v->cnt_sorted.as<int>();

    long long* stats_dev = v->stats.as<long long>();
    long long* stats_h = v->pinned.ptr;

    if ( diag )
    {
        CB_CUDA( cudaMemsetAsync( v->ctrl.ptr, 0, 64, stream ) );
        a.mask_capacity = 1ll << 40;
        if ( n > 0 )
            CB_TRY( tile_diag_pass( a, half, stream ) );
        CB_CUDA( cudaMemcpyAsync( stats_h, v->ctrl.ptr, 32, cudaMemcpyDeviceToHost, stream ) );
        CB_CUDA( cudaStreamSynchronize( stream ) );
        float e;
        unsigned misses, flips;
        memcpy( &e, reinterpret_cast<char*>( stats_h ) + 16, 4 );
        memcpy( &misses, reinterpret_cast<char*>( stats_h ) + 20, 4 );
        memcpy( &flips, reinterpret_cast<char*>( stats_h ) + 24, 4 );
        diag[0] = (double)e;
        diag[2] = (double)misses; // pairs with s <= 4 r^2 whose error exceeded the bound
        diag[3] = (double)flips;  // values outside +-tau with the wrong sign (any distance)
        return CB_OK;
    }

    // first guess of the mask buffer (~100 B per particle at liquid densities); the count
    // pass reports the exact need if it does not fit
    if ( v->masks.capacity == 0 )
        CB_TRY( v->masks.ensure( (size_t)( 128.0 * (double)n ) + ( 1u << 20 ) ) );
    // One host wait per build, and it is off the GPU's critical path: when the list of a
    // previous build left a neighbour array behind (the steady state of a rebuild loop), the CSR
    // fill pass is launched BEFORE the host has seen the new size -- the kernel itself checks
    // that the masks did not overflow and that the list fits, and leaves without writing
    // otherwise -- so the host wakes up while the fill pass is already running.  Only when the
    // speculation fails (first build, list grew past the 5 % slack) does the fill pass start
    // after the read-back, as the reference's processCounts does (:507-562).
    if ( !v->ev_stats )
        CB_CUDA( cudaEventCreateWithFlags( &v->ev_stats, cudaEventDisableTiming ) );
    bool filled = false;
    for ( int attempt = 0;; ++attempt )
    {
        a.masks = v->masks.as<uint4>();
        a.mask_capacity = (long long)( v->masks.capacity / ( kChunkWords * 4 ) );
        CB_CUDA( cudaMemsetAsync( v->ctrl.ptr, 0, 64, stream ) );
        CB_CUDA( cudaMemsetAsync( v->cnt_sorted.ptr, 0, sizeof( int ) * (size_t)( ns_cap + 8 ),
                                  stream ) );
        if ( n > 0 )
            CB_TRY( tile_count_pass( a, half, stream ) );
        v->mark( 3, stream );
        if ( !csr )
            CB_TRY( max_and_sum_i32( v->counts.as<int>(), n, stats_dev, stream ) );
        if ( csr )
        {
            // offsets + max + sum of the counts in ONE pass over them
            CB_TRY( exclusive_scan_stats_i32( v->counts.as<int>(), v->offsets.as<int>(), n,
                                              stats_dev, v->scan, stream ) );
            a.offsets = v->offsets.as<int>();
            if ( n > 0 )
                CB_TRY( tile_sorted_dst( a, tg.ncells, ns_cap, v->dst_sorted.as<int>(), stream ) );
            a.dst_sorted = v->dst_sorted.as<int>();
        }
        CB_CUDA( cudaMemcpyAsync( stats_h, stats_dev, 2 * sizeof( long long ),
                                  cudaMemcpyDeviceToHost, stream ) );
        CB_CUDA( cudaMemcpyAsync( stats_h + 2, v->ctrl.ptr, 2 * sizeof( long long ),
                                  cudaMemcpyDeviceToHost, stream ) );
        CB_CUDA( cudaMemcpyAsync( stats_h + 4, v->chunk_off.as<int>() + rec_capacity,
                                  sizeof( int ), cudaMemcpyDeviceToHost, stream ) );
        if ( n_dev )
            CB_CUDA( cudaMemcpyAsync( stats_h + 5, n_dev, sizeof( long long ),
                                      cudaMemcpyDeviceToHost, stream ) );
        v->mark( 4, stream );
        CB_CUDA( cudaEventRecord( v->ev_stats, stream ) );
        const char* spec_env = getenv( "CB_VERLET_SPECULATE" );
        const bool speculate = csr && attempt == 0 && n > 0 &&
                               v->neighbors.capacity >= sizeof( int ) &&
                               !( spec_env && spec_env[0] == '0' );
        if ( speculate )
        {
            a.neighbors = v->neighbors.as<int>();
            a.width = 0;
            a.spec_total = stats_dev + 1;
            a.spec_capacity = (long long)( v->neighbors.capacity / sizeof( int ) );
            a.spec_failed = reinterpret_cast<int*>( v->ctrl.as<char>() + 32 );
            CB_TRY( tile_fill_pass( a, true, stream ) );
            a.spec_total = nullptr;
        }
        CB_CUDA( cudaEventSynchronize( v->ev_stats ) ); // sizes the allocation (:518-528)
        const int overflowed = (int)( stats_h[3] & 0xffffffffll );
        if ( !overflowed )
        {
            filled = speculate && stats_h[1] <= (long long)( v->neighbors.capacity / sizeof( int ) );
            break;
        }
        if ( attempt >= 1 )
            return fail( CB_ERR_NOMEM, "cb_verlet_build: mask buffer kept overflowing" );
        const long long chunks = (long long)( stats_h[4] & 0xffffffffll );
        CB_TRY( v->masks.ensure( (size_t)( chunks + 16 ) * kChunkWords * 4, 1.05 ) );
        CB_CUDA( cudaMemsetAsync( v->counts.ptr, 0, sizeof( int ) * na, stream ) );
    }
    v->max_n = stats_h[0];
    v->total = stats_h[1];
    v->extent = v->total;
    if ( n_dev )
        v->n = stats_h[5] < n ? stats_h[5] : n;
    if ( csr )
    {
        if ( v->total > 2147483647ll )
            return fail( CB_ERR_OVERFLOW, "cb_verlet_build: total neighbours exceed INT_MAX" );
        if ( !filled )
            CB_TRY( v->neighbors.ensure( sizeof( int ) * (size_t)( v->total > 0 ? v->total : 1 ),
                                         1.05 ) );
        a.offsets = v->offsets.as<int>();
        a.width = 0;
    }
    else
    {
        // initCounts / processCounts(2D) (:495-505, :536-562): keep max_neigh when it is
        // enough, otherwise reallocate to exactly max_n ("refill")
        if ( max_neigh > 0 && v->max_n <= max_neigh )
            v->width = max_neigh;
        else
        {
            v->width = v->max_n;
            if ( max_neigh > 0 )
                v->refilled = 1;
        }
        if ( (double)na * (double)v->width > 9.0e18 )
            return fail( CB_ERR_OVERFLOW, "cb_verlet_build: 2D list too large" );
        CB_TRY( v->neighbors.ensure( sizeof( int ) * na *
                                     (size_t)( v->width > 0 ? v->width : 1 ) ) );
        a.offsets = nullptr;
        a.width = v->width;
    }
    a.neighbors = v->neighbors.as<int>();
    if ( !filled && n > 0 && v->total > 0 )
        CB_TRY( tile_fill_pass( a, csr, stream ) );
    v->mark( 5, stream );
    v->built = true;
    return CB_OK;
}

extern "C" int cb_verlet_build( cb_verlet* v, const cb_positions* x, int64_t begin,
                                int64_t end, double radius, double cell_size_ratio,
                                const double* grid_min, const double* grid_max,
                                int64_t max_neigh, int algorithm, int layout,
                                int build_op, cb_stream_t stream_ )
{
    return cb::verlet_build_devcount( v, x, begin, end, radius, cell_size_ratio, grid_min,
                                      grid_max, max_neigh, algorithm, layout, build_op, stream_,
                                      nullptr );
}

bool cb::verlet_devcount_supported()
{
    const char* impl_env = getenv( "CB_VERLET_IMPL" );
    return !( impl_env && ( strcmp( impl_env, "v0" ) == 0 || strcmp( impl_env, "v1" ) == 0 ) );
}

int cb::verlet_build_devcount( cb_verlet* v, const cb_positions* x, long long begin,
                               long long end, double radius, double cell_size_ratio,
                               const double* grid_min, const double* grid_max,
                               long long max_neigh, int algorithm, int layout, int build_op,
                               void* stream_, const long long* n_dev )
{
    ScopedRegion region( "Cabana::VerletList::build" );
    if ( n_dev && !verlet_devcount_supported() )
        return fail( CB_ERR_UNSUPPORTED,
                     "cb_verlet_build: a device-side particle count needs the v2 kernels" );
    if ( !v || !x || !grid_min || !grid_max )
        return fail( CB_ERR_INVALID, "cb_verlet_build: null argument" );
    // asserts of the reference (:1381-1382)
    if ( begin < 0 || end < begin || end > x->n )
        return fail( CB_ERR_INVALID, "cb_verlet_build: bad particle range" );
    if ( !( radius > 0.0 ) || !( cell_size_ratio > 0.0 ) )
        return fail( CB_ERR_INVALID, "cb_verlet_build: radius and ratio must be > 0" );
    if ( x->vlen < 1 || max_neigh < 0 )
        return fail( CB_ERR_INVALID, "cb_verlet_build: bad descriptor" );
    if ( algorithm != CB_NEIGHBOR_FULL && algorithm != CB_NEIGHBOR_HALF )
        return fail( CB_ERR_INVALID, "cb_verlet_build: bad algorithm tag" );
    if ( layout != CB_LAYOUT_CSR && layout != CB_LAYOUT_2D )
        return fail( CB_ERR_INVALID, "cb_verlet_build: bad layout tag" );
    // TeamOpTag and TeamVectorOpTag differ only in how Kokkos maps the candidate loop
    // (:418-442); the result is the same and so is the kernel here.
    if ( build_op != CB_OP_TEAM && build_op != CB_OP_TEAM_VECTOR )
        return fail( CB_ERR_INVALID, "cb_verlet_build: bad build tag" );
    if ( x->n >= 2147483647ll )
        return fail( CB_ERR_UNSUPPORTED, "cb_verlet_build: ids are 32-bit int (:57-63)" );
    for ( int d = 0; d < 3; ++d )
        if ( !( grid_max[d] > grid_min[d] ) )
            return fail( CB_ERR_INVALID, "cb_verlet_build: grid_max <= grid_min" );

    cudaStream_t stream = (cudaStream_t)stream_;
    const long long n = x->n;
    const size_t na = (size_t)( n > 0 ? n : 1 );

    // init (:205-240): delta = cell_size_ratio * radius for the binning grid (:224);
    // the stencil grid uses radius * cell_size_ratio (Cabana_LinkedCellList.hpp:60) --
    // the same double -- so one grid serves both.
    const double grid_size = cell_size_ratio * radius;
    const double delta[3] = { grid_size, grid_size, grid_size };
    cb_grid grid;
    cb_grid_init( &grid, grid_min, grid_max, delta );
    const long long ncell = (long long)grid.nx[0] * grid.nx[1] * grid.nx[2];
    if ( grid.nx[0] <= 0 || grid.nx[1] <= 0 || grid.nx[2] <= 0 ||
         ncell >= 2147483647ll )
        return fail( CB_ERR_INVALID, "cb_verlet_build: cell count out of int range" );
    const int cell_range = cb_stencil_cell_range( cell_size_ratio );
    const double rsqr = radius * radius; // :239

    // Kernel generation: v2 (tile kernels, default), v1 (refined grid + FP32 SIMT filter,
    // CB_VERLET_IMPL=v1) or v0 (reference-shaped exact FP64, CB_VERLET_IMPL=v0).
    const char* impl_env = getenv( "CB_VERLET_IMPL" );
    const bool use_v0 = impl_env && strcmp( impl_env, "v0" ) == 0;
    const bool use_v1 = impl_env && strcmp( impl_env, "v1" ) == 0;
    if ( !use_v0 && !use_v1 )
    {
        v->built = false;
        v->layout = layout;
        v->algorithm = algorithm;
        v->n = n;
        v->total = 0;
        v->max_n = 0;
        v->width = 0;
        v->refilled = 0;
        return build_tile( v, x, begin, end, radius, grid, cell_range, grid_min, grid_max,
                           max_neigh, algorithm, layout, stream, nullptr, n_dev );
    }
    int refine = 1;
    if ( !use_v0 )
    {
        // refined cells of about r/2; m is a power of two (cheap user-cell arithmetic)
        const double target = 2.0 * grid.dx[0] / radius;
        refine = 1;
        while ( refine < 8 && (double)refine * 1.4142 < target )
            refine *= 2;
        const char* m_env = getenv( "CB_VERLET_REFINE" );
        if ( m_env && ( atoi( m_env ) == 1 || atoi( m_env ) == 2 || atoi( m_env ) == 4 ||
                        atoi( m_env ) == 8 ) )
            refine = atoi( m_env );
        // keep the refined grid a sane size relative to the particle count
        while ( refine > 1 )
        {
            const double nc = (double)ncell * refine * refine * refine;
            if ( nc < 2.0e9 && nc <= 8.0 * (double)n + 1048576.0 )
                break;
            refine /= 2;
        }
    }
    // v1 tabulates the z reach of a stencil row for |da|,|db| <= 8 refined cells; finer
    // user grids (cell_size_ratio < ~0.13) take the v0 kernels.
    bool fine_ok = !use_v0;
    if ( fine_ok )
        for ( int d = 0; d < 3; ++d )
        {
            const double hf = grid.dx[d] / refine;
            if ( floor( radius * ( 1.0 + 1.0e-9 ) / hf ) + 1.0 > 8.0 )
                fine_ok = false;
        }
    const bool use_fine = fine_ok;
    if ( !use_fine )
        refine = 1;
    const long long ncell_f = ncell * refine * refine * refine;

    CB_TRY( v->pinned.ensure() );
    CB_TRY( v->stats.ensure( 4 * sizeof( long long ) ) );
    CB_TRY( v->counts.ensure( sizeof( int ) * na, 1.1 ) );
    CB_TRY( v->cell_counts.ensure( sizeof( int ) * (size_t)ncell_f ) );
    CB_TRY( v->cell_off.ensure( sizeof( unsigned ) * (size_t)( ncell_f + 1 ) ) );
    const char* col_env = getenv( "CB_VERLET_COLUMNS" );
    const bool use_columns = use_fine && !( col_env && strcmp( col_env, "0" ) == 0 );
    if ( use_fine )
        CB_TRY( v->q.ensure( sizeof( float4 ) * ( na + 1 ), 1.1 ) );
    if ( use_columns )
        CB_TRY( v->worklist.ensure( sizeof( unsigned ) * (size_t)( ncell_f + 4 ) ) );
    CB_TRY( v->permute.ensure( sizeof( unsigned ) * na, 1.1 ) );
    CB_TRY( v->cell_of.ensure( sizeof( int ) * na, 1.1 ) );
    CB_TRY( v->xs.ensure( sizeof( double ) * na, 1.1 ) );
    CB_TRY( v->ys.ensure( sizeof( double ) * na, 1.1 ) );
    CB_TRY( v->zs.ensure( sizeof( double ) * na, 1.1 ) );

    v->built = false;
    v->layout = layout;
    v->algorithm = algorithm;
    v->n = n;
    v->total = 0;
    v->max_n = 0;
    v->width = 0;
    v->refilled = 0;

    v->begin_profile();
    v->mark( 0, stream );

    // counts zero-initialised (:215-216)
    CB_CUDA( cudaMemsetAsync( v->counts.ptr, 0, sizeof( int ) * na, stream ) );

    // LCL over ALL particles, not just [begin,end) (:229-235).
    CB_TRY( bin_particles( grid, *x, 0, n, v->cell_counts.as<int>(),
                           v->cell_off.as<unsigned>(), v->permute.as<unsigned>(),
                           v->cell_of.as<int>(), v->rank, v->scan, stream, refine ) );
    v->mark( 1, stream );
    if ( n > 0 )
    {
        k_gather_sorted<<<launch_grid_for( n, kBlock ), kBlock, 0, stream>>>(
            make_access( *x ), n, v->permute.as<unsigned>(), v->xs.as<double>(),
            v->ys.as<double>(), v->zs.as<double>(), use_fine ? v->q.as<float4>() : nullptr,
            grid.min[0], grid.min[1], grid.min[2] );
        CB_CHECK_LAUNCH();
    }
    v->mark( 2, stream );

    VerletArgs a;
    a.xs = v->xs.as<double>();
    a.ys = v->ys.as<double>();
    a.zs = v->zs.as<double>();
    a.ids = v->permute.as<unsigned>();
    a.cell_off = v->cell_off.as<unsigned>();
    a.g = to_grid( grid );
    a.cell_range = cell_range;
    a.rsqr = rsqr;
    a.n = n;
    a.begin = begin;
    a.end = end;
    a.counts = v->counts.as<int>();
    a.offsets = nullptr;
    a.neighbors = nullptr;
    a.width = 0;

    FineArgs fa;
    memset( &fa, 0, sizeof( fa ) );
    if ( use_fine )
    {
        fa.q = v->q.as<float4>();
        fa.ids = a.ids;
        fa.xs = a.xs;
        fa.ys = a.ys;
        fa.zs = a.zs;
        fa.cell_off = a.cell_off;
        fa.ug = a.g;
        fa.lgm = refine == 8 ? 3 : ( refine == 4 ? 2 : ( refine == 2 ? 1 : 0 ) );
        fa.R = cell_range;
        double extent = 0.0, cmax = 0.0, hmax = 0.0;
        for ( int d = 0; d < 3; ++d )
        {
            fa.nf[d] = grid.nx[d] * refine;
            fa.hf[d] = grid.dx[d] / refine;
            extent = fmax( extent, grid.max[d] - grid.min[d] );
            cmax = fmax( cmax, fmax( fabs( grid.min[d] ), fabs( grid.max[d] ) ) );
            hmax = fmax( hmax, fa.hf[d] );
        }
        // Conservative reach for cell-level pruning: a point's refined cell is known to
        // ~1e-12 h, so two points |k| cells apart are at least (|k|-1) h - eps apart.
        const double reach = radius * ( 1.0 + 1.0e-9 ) + 4.0e-9 * hmax;
        const double reach2 = reach * reach;
        for ( int d = 0; d < 3; ++d )
        {
            fa.K[d] = (int)floor( reach / fa.hf[d] ) + 1;
            const int kmax = ( cell_range + 1 ) * refine; // never beyond the user stencil
            if ( fa.K[d] > kmax )
                fa.K[d] = kmax;
            if ( fa.K[d] > 8 )
                fa.K[d] = 8; // cannot bind: use_fine guarantees floor(reach/hf)+1 <= 8
        }
        for ( int da = 0; da <= 8; ++da )
            for ( int db = 0; db <= 8; ++db )
            {
                const double ga = ( da > 1 ? da - 1 : 0 ) * fa.hf[0];
                const double gb = ( db > 1 ? db - 1 : 0 ) * fa.hf[1];
                const double rem = reach2 - ( ga * ga + gb * gb );
                int kz = -1;
                if ( rem >= 0.0 && da <= fa.K[0] && db <= fa.K[1] )
                {
                    kz = fa.K[2];
                    while ( kz > 1 )
                    {
                        const double gz = ( kz - 1 ) * fa.hf[2];
                        if ( gz * gz <= rem )
                            break;
                        --kz;
                    }
                }
                fa.kz[da * 9 + db] = (signed char)kz;
            }
        fa.rsqr = rsqr;
        // Tier-1 filter: |s32 - s| <= E for s <= 1.1 r^2 with origin-relative floats
        // (DESIGN.md "Exactness"); tau = 4 E + 2^-22 r^2.
        const double u = ldexp( 1.0, -24 );
        const double E = u * ( 6.5 * 1.1 * rsqr + 7.3 * 1.05 * extent * radius );
        const double tau = 4.0 * E + ldexp( rsqr, -22 );
        fa.t_hi = nextafterf( (float)( rsqr + tau ), INFINITY );
        fa.t_lo = rsqr - tau > 0.0 ? nextafterf( (float)( rsqr - tau ), -INFINITY ) : -1.0f;
        // Band in which the reference's cell prune can reject an in-range pair: its min
        // distance is computed with O(ulp(coordinate)) error.
        const double eta = ldexp( cmax + extent + radius, -46 );
        fa.band = 8.0 * radius * eta + 4.0 * eta * eta + ldexp( rsqr, -44 );
        fa.n = n;
        fa.begin = begin;
        fa.end = end;
        fa.full_range = ( begin == 0 && end == n ) ? 1 : 0;
        fa.ncell = ncell_f;
        fa.counts = a.counts;
        fa.offsets = nullptr;
        fa.neighbors = nullptr;
        fa.width = 0;
        fa.worklist = use_columns ? v->worklist.as<unsigned>() + 4 : nullptr;
        fa.work_count = use_columns ? v->worklist.as<unsigned>() : nullptr;
    }
    auto run_pass = [&]( bool fill ) -> int
    {
        return fill ? launch_pass<kFill>( a, algorithm, layout, stream )
                    : launch_pass<kCount>( a, algorithm, layout, stream );
    };

    long long* stats_dev = v->stats.as<long long>();
    long long* stats_h = v->pinned.ptr;

    if ( use_fine )
    {
        // ---- v1: ONE test pass into a binned-order temporary, then reorder ----------
        CB_TRY( v->tmp_off.ensure( sizeof( unsigned ) * na, 1.1 ) );
        CB_TRY( v->ctrl.ensure( 64 ) );
        if ( layout == CB_LAYOUT_CSR )
            CB_TRY( v->offsets.ensure( sizeof( int ) * ( na + 1 ), 1.1 ) );
        // first guess of the temporary's size: uniform-density estimate plus the slack
        // of the per-warp reservations; a rebuild reuses the previous capacity
        if ( v->tmp.capacity == 0 )
        {
            double vol = 1.0;
            for ( int d = 0; d < 3; ++d )
                vol *= grid.max[d] - grid.min[d];
            double k_est = 4.18879 * radius * radius * radius * (double)n / vol *
                           ( algorithm == CB_NEIGHBOR_HALF ? 0.5 : 1.0 );
            if ( k_est > (double)( n > 1 ? n - 1 : 1 ) )
                k_est = (double)( n > 1 ? n - 1 : 1 ); // a row never holds more than n - 1 ids
            // + the slack of the per-warp reservations of a persistent grid (a few thousand
            // warps x a few hundred ids); the pass reports the exact need if this is short
            double est = 1.15 * k_est * (double)n + 16.0 * (double)n + 4.0e6;
            if ( v->tmp.ensure( sizeof( int ) * (size_t)est ) != CB_OK )
            {
                // too optimistic for this device: start small and let the overflow retry size it
                (void)cudaGetLastError();
                est = 16.0 * (double)n + 4.0e6;
                CB_TRY( v->tmp.ensure( sizeof( int ) * (size_t)est ) );
            }
        }
        fa.tmp_off = v->tmp_off.as<unsigned>();
        const bool binned_rows = v->row_placement == CB_ROWS_BINNED && layout == CB_LAYOUT_CSR;
        fa.offsets_direct = binned_rows ? v->offsets.as<int>() : nullptr;
        fa.cursor = reinterpret_cast<unsigned long long*>( v->ctrl.as<char>() );
        fa.overflow = reinterpret_cast<int*>( v->ctrl.as<char>() + 8 );
        for ( int attempt = 0;; ++attempt )
        {
            fa.tmp = v->tmp.as<int>();
            fa.tmp_capacity = (long long)( v->tmp.capacity / sizeof( int ) );
            CB_CUDA( cudaMemsetAsync( v->ctrl.ptr, 0, 64, stream ) );
            CB_TRY( launch_fine_single( fa, algorithm, stream ) );
            v->mark( 3, stream );
            CB_TRY( max_and_sum_i32( v->counts.as<int>(), n, stats_dev, stream ) );
            if ( layout == CB_LAYOUT_CSR && !binned_rows )
                CB_TRY( exclusive_scan_i32( v->counts.as<int>(), v->offsets.as<int>(), n,
                                            false, nullptr, v->scan, stream ) );
            CB_CUDA( cudaMemcpyAsync( stats_h, stats_dev, 2 * sizeof( long long ),
                                      cudaMemcpyDeviceToHost, stream ) );
            CB_CUDA( cudaMemcpyAsync( stats_h + 2, v->ctrl.ptr, 2 * sizeof( long long ),
                                      cudaMemcpyDeviceToHost, stream ) );
            v->mark( 4, stream );
            CB_CUDA( cudaStreamSynchronize( stream ) ); // sizes the allocation (:518-528)
            const int overflowed = (int)( stats_h[3] & 0xffffffffll );
            // row starts inside the temporary are 32-bit
            if ( stats_h[2] > 4294967295ll )
                return fail( CB_ERR_OVERFLOW,
                             "cb_verlet_build: more than 2^32 stored ids (v1 temporary)" );
            if ( !overflowed )
                break;
            if ( attempt >= 2 )
                return fail( CB_ERR_NOMEM, "cb_verlet_build: temporary buffer kept overflowing" );
            CB_TRY( v->tmp.ensure( sizeof( int ) * (size_t)( (double)stats_h[2] * 1.05 + 1.0e6 ) ) );
        }
        v->max_n = stats_h[0];
        v->total = stats_h[1];
        if ( binned_rows )
        {
            // the temporary buffer IS the neighbour array: offsets[pid] were written by the
            // test pass; rows outside [begin,end) keep count 0 (their offsets are unused)
            if ( stats_h[2] > 2147483647ll )
                return fail( CB_ERR_OVERFLOW,
                             "cb_verlet_build: neighbour array exceeds INT_MAX entries" );
            v->neighbors.swap( v->tmp );
            v->extent = stats_h[2] > 0 ? stats_h[2] : 0;
            v->width = 0;
            v->mark( 5, stream );
            v->built = true;
            return CB_OK;
        }
        v->extent = v->total;
        if ( layout == CB_LAYOUT_CSR )
        {
            if ( v->total > 2147483647ll )
                return fail( CB_ERR_OVERFLOW,
                             "cb_verlet_build: total neighbours exceed INT_MAX" );
            CB_TRY( v->neighbors.ensure(
                sizeof( int ) * (size_t)( v->total > 0 ? v->total : 1 ), 1.05 ) );
            fa.offsets = v->offsets.as<int>();
            fa.width = 0;
        }
        else
        {
            // initCounts / processCounts(2D) (:495-505, :536-562): keep max_neigh when it is
            // enough, otherwise reallocate to exactly max_n ("refill")
            if ( max_neigh > 0 && v->max_n <= max_neigh )
                v->width = max_neigh;
            else
            {
                v->width = v->max_n;
                if ( max_neigh > 0 )
                    v->refilled = 1;
            }
            if ( (double)na * (double)v->width > 9.0e18 )
                return fail( CB_ERR_OVERFLOW, "cb_verlet_build: 2D list too large" );
            CB_TRY( v->neighbors.ensure(
                sizeof( int ) * na * (size_t)( v->width > 0 ? v->width : 1 ) ) );
            fa.offsets = nullptr;
            fa.width = v->width;
        }
        fa.neighbors = v->neighbors.as<int>();
        CB_TRY( launch_fine_reorder( fa, layout, stream ) );
        v->mark( 5, stream );
        v->built = true;
        return CB_OK;
    }

    if ( layout == CB_LAYOUT_CSR )
    {
        // count -> processCounts(CSR) -> fill  (:1454-1480, :507-532)
        CB_TRY( run_pass( false ) );
        v->mark( 3, stream );
        CB_TRY( v->offsets.ensure( sizeof( int ) * ( na + 1 ), 1.1 ) );
        CB_TRY( max_and_sum_i32( v->counts.as<int>(), n, stats_dev, stream ) );
        CB_TRY( exclusive_scan_i32( v->counts.as<int>(), v->offsets.as<int>(), n, false,
                                    nullptr, v->scan, stream ) );
        CB_CUDA( cudaMemcpyAsync( stats_h, stats_dev, 2 * sizeof( long long ),
                                  cudaMemcpyDeviceToHost, stream ) );
        v->mark( 4, stream );
        CB_CUDA( cudaStreamSynchronize( stream ) ); // sizes the allocation (:518-528)
        v->max_n = stats_h[0];
        v->total = stats_h[1];
        v->extent = v->total;
        if ( v->total > 2147483647ll )
            return fail( CB_ERR_OVERFLOW,
                         "cb_verlet_build: total neighbours exceed INT_MAX" );
        CB_TRY( v->neighbors.ensure( sizeof( int ) * (size_t)( v->total > 0 ? v->total : 1 ),
                                     1.05 ) );
        a.offsets = v->offsets.as<int>();
        a.neighbors = v->neighbors.as<int>();
        CB_TRY( run_pass( true ) );
        v->mark( 5, stream );
    }
    else
    {
        bool count = true;
        if ( max_neigh > 0 ) // initCounts(2D) :495-505
        {
            count = false;
            v->width = max_neigh;
            CB_TRY( v->neighbors.ensure( sizeof( int ) * na * (size_t)v->width ) );
            a.neighbors = v->neighbors.as<int>();
            a.width = v->width;
            CB_TRY( run_pass( true ) );
        }
        else
        {
            CB_TRY( run_pass( false ) );
        }
        v->mark( 3, stream );
        // processCounts(2D) :536-562
        CB_TRY( max_and_sum_i32( v->counts.as<int>(), n, stats_dev, stream ) );
        CB_CUDA( cudaMemcpyAsync( stats_h, stats_dev, 2 * sizeof( long long ),
                                  cudaMemcpyDeviceToHost, stream ) );
        v->mark( 4, stream );
        CB_CUDA( cudaStreamSynchronize( stream ) );
        v->max_n = stats_h[0];
        v->total = stats_h[1];
        v->extent = v->total;
        if ( count || v->max_n > v->width )
        {
            v->width = v->max_n;
            if ( (double)na * (double)v->width > 9.0e18 )
                return fail( CB_ERR_OVERFLOW, "cb_verlet_build: 2D list too large" );
            CB_TRY( v->neighbors.ensure(
                sizeof( int ) * na * (size_t)( v->width > 0 ? v->width : 1 ) ) );
            a.neighbors = v->neighbors.as<int>();
            a.width = v->width;
            CB_TRY( run_pass( true ) );
            if ( !count )
                v->refilled = 1;
        }
        v->mark( 5, stream );
    }
    v->built = true;
    return CB_OK;
}

extern "C" int cb_verlet_build_radii( cb_verlet* v, const cb_positions* x,
                                      const cb_field* radii, int64_t begin, int64_t end,
                                      double background_radius, double cell_size_ratio,
                                      const double* grid_min, const double* grid_max,
                                      int64_t max_neigh, int algorithm, int layout,
                                      int build_op, cb_stream_t stream_ )
{
    ScopedRegion region( "Cabana::VerletList::build" );
    if ( !v || !x || !radii || !grid_min || !grid_max )
        return fail( CB_ERR_INVALID, "cb_verlet_build_radii: null argument" );
    if ( begin < 0 || end < begin || end > x->n )
        return fail( CB_ERR_INVALID, "cb_verlet_build_radii: bad particle range" );
    if ( !( background_radius > 0.0 ) || !( cell_size_ratio > 0.0 ) || x->vlen < 1 ||
         max_neigh < 0 )
        return fail( CB_ERR_INVALID, "cb_verlet_build_radii: bad argument" );
    // assert( size( positions ) == size( neighborhood_radius ) ) (:194)
    if ( radii->n != x->n || radii->elem_bytes != 8 || radii->num_comp != 1 || radii->vlen < 1 )
        return fail( CB_ERR_INVALID,
                     "cb_verlet_build_radii: radii must be one double per particle" );
    if ( ( algorithm != CB_NEIGHBOR_FULL && algorithm != CB_NEIGHBOR_HALF ) ||
         ( layout != CB_LAYOUT_CSR && layout != CB_LAYOUT_2D ) ||
         ( build_op != CB_OP_TEAM && build_op != CB_OP_TEAM_VECTOR ) )
        return fail( CB_ERR_INVALID, "cb_verlet_build_radii: bad tag" );
    if ( x->n >= 2147483647ll )
        return fail( CB_ERR_UNSUPPORTED, "cb_verlet_build_radii: ids are 32-bit int" );
    for ( int d = 0; d < 3; ++d )
        if ( !( grid_max[d] > grid_min[d] ) )
            return fail( CB_ERR_INVALID, "cb_verlet_build_radii: grid_max <= grid_min" );

    cudaStream_t stream = (cudaStream_t)stream_;
    const long long n = x->n;
    const size_t na = (size_t)( n > 0 ? n : 1 );
    const double grid_size = cell_size_ratio * background_radius; // :224
    const double delta[3] = { grid_size, grid_size, grid_size };
    cb_grid grid;
    cb_grid_init( &grid, grid_min, grid_max, delta );
    const long long ncell = (long long)grid.nx[0] * grid.nx[1] * grid.nx[2];
    if ( grid.nx[0] <= 0 || grid.nx[1] <= 0 || grid.nx[2] <= 0 || ncell >= 2147483647ll )
        return fail( CB_ERR_INVALID, "cb_verlet_build_radii: cell count out of int range" );

    CB_TRY( v->pinned.ensure() );
    CB_TRY( v->stats.ensure( 4 * sizeof( long long ) ) );
    CB_TRY( v->counts.ensure( sizeof( int ) * na, 1.1 ) );
    CB_TRY( v->row_cursor.ensure( sizeof( int ) * na, 1.1 ) );
    CB_TRY( v->cell_counts.ensure( sizeof( int ) * (size_t)ncell ) );
    CB_TRY( v->cell_off.ensure( sizeof( unsigned ) * (size_t)( ncell + 1 ) ) );
    CB_TRY( v->permute.ensure( sizeof( unsigned ) * na, 1.1 ) );
    CB_TRY( v->cell_of.ensure( sizeof( int ) * na, 1.1 ) );
    CB_TRY( v->xs.ensure( sizeof( double ) * na, 1.1 ) );
    CB_TRY( v->ys.ensure( sizeof( double ) * na, 1.1 ) );
    CB_TRY( v->zs.ensure( sizeof( double ) * na, 1.1 ) );
    v->built = false;
    v->layout = layout;
    v->algorithm = algorithm;
    v->n = n;
    v->total = v->max_n = v->width = 0;
    v->refilled = 0;
    v->begin_profile();

    CB_CUDA( cudaMemsetAsync( v->counts.ptr, 0, sizeof( int ) * na, stream ) );
    CB_CUDA( cudaMemsetAsync( v->row_cursor.ptr, 0, sizeof( int ) * na, stream ) );
    CB_TRY( bin_particles( grid, *x, 0, n, v->cell_counts.as<int>(), v->cell_off.as<unsigned>(),
                           v->permute.as<unsigned>(), v->cell_of.as<int>(), v->rank, v->scan,
                           stream, 1 ) );
    if ( n > 0 )
    {
        k_gather_sorted<<<launch_grid_for( n, kBlock ), kBlock, 0, stream>>>(
            make_access( *x ), n, v->permute.as<unsigned>(), v->xs.as<double>(),
            v->ys.as<double>(), v->zs.as<double>(), nullptr, grid.min[0], grid.min[1],
            grid.min[2] );
        CB_CHECK_LAUNCH();
    }
    RadiiArgs ra;
    VerletArgs& a = ra.v;
    a.xs = v->xs.as<double>();
    a.ys = v->ys.as<double>();
    a.zs = v->zs.as<double>();
    a.ids = v->permute.as<unsigned>();
    a.cell_off = v->cell_off.as<unsigned>();
    a.g = to_grid( grid );
    a.cell_range = cb_stencil_cell_range( cell_size_ratio );
    a.rsqr = background_radius * background_radius;
    a.n = n;
    a.begin = begin;
    a.end = end;
    a.counts = v->counts.as<int>();
    a.offsets = nullptr;
    a.neighbors = nullptr;
    a.width = 0;
    ra.radii = make_access( *radii );
    ra.cursor = v->row_cursor.as<int>();

    CB_TRY( launch_radii_pass<kCount>( ra, algorithm, layout, stream ) );
    long long* stats_dev = v->stats.as<long long>();
    long long* stats_h = v->pinned.ptr;
    CB_TRY( max_and_sum_i32( v->counts.as<int>(), n, stats_dev, stream ) );
    if ( layout == CB_LAYOUT_CSR )
    {
        CB_TRY( v->offsets.ensure( sizeof( int ) * ( na + 1 ), 1.1 ) );
        CB_TRY( exclusive_scan_i32( v->counts.as<int>(), v->offsets.as<int>(), n, false, nullptr,
                                    v->scan, stream ) );
    }
    CB_CUDA( cudaMemcpyAsync( stats_h, stats_dev, 2 * sizeof( long long ), cudaMemcpyDeviceToHost,
                              stream ) );
    CB_CUDA( cudaStreamSynchronize( stream ) );
    v->max_n = stats_h[0];
    v->total = stats_h[1];
    v->extent = v->total;
    if ( layout == CB_LAYOUT_CSR )
    {
        if ( v->total > 2147483647ll )
            return fail( CB_ERR_OVERFLOW, "cb_verlet_build_radii: total neighbours exceed INT_MAX" );
        CB_TRY( v->neighbors.ensure( sizeof( int ) * (size_t)( v->total > 0 ? v->total : 1 ),
                                     1.05 ) );
        a.offsets = v->offsets.as<int>();
    }
    else
    {
        if ( max_neigh > 0 && v->max_n <= max_neigh )
            v->width = max_neigh;
        else
        {
            v->width = v->max_n;
            if ( max_neigh > 0 )
                v->refilled = 1;
        }
        if ( (double)na * (double)v->width > 9.0e18 )
            return fail( CB_ERR_OVERFLOW, "cb_verlet_build_radii: 2D list too large" );
        CB_TRY( v->neighbors.ensure( sizeof( int ) * na *
                                     (size_t)( v->width > 0 ? v->width : 1 ) ) );
        a.width = v->width;
    }
    a.neighbors = v->neighbors.as<int>();
    if ( v->total > 0 )
        CB_TRY( launch_radii_pass<kFill>( ra, algorithm, layout, stream ) );
    v->built = true;
    return CB_OK;
}

extern "C" int cb_verlet_filter_selftest( cb_verlet* v, const cb_positions* x, double radius,
                                          const double* grid_min, const double* grid_max,
                                          int algorithm, double* out_h, cb_stream_t stream_ )
{
    if ( !v || !x || !grid_min || !grid_max || !out_h )
        return fail( CB_ERR_INVALID, "cb_verlet_filter_selftest: null argument" );
    if ( !( radius > 0.0 ) || x->vlen < 1 || x->n >= 2147483647ll )
        return fail( CB_ERR_INVALID, "cb_verlet_filter_selftest: bad argument" );
    const double delta[3] = { radius, radius, radius };
    cb_grid grid;
    cb_grid_init( &grid, grid_min, grid_max, delta );
    v->built = false;
    return build_tile( v, x, 0, x->n, radius, grid, 1, grid_min, grid_max, 0, algorithm,
                       CB_LAYOUT_CSR, (cudaStream_t)stream_, out_h );
}

extern "C" int cb_verlet_set_profiling( cb_verlet* v, int enable )
{
    if ( !v )
        return fail( CB_ERR_INVALID, "cb_verlet_set_profiling: null argument" );
    if ( enable && !v->have_events )
    {
        for ( auto& set : v->ev )
            for ( auto& e : set )
                CB_CUDA( cudaEventCreate( &e ) );
        v->have_events = true;
    }
    v->profiling = enable != 0;
    v->reset_profile(); // the averages of cb_verlet_get_phase_times start here
    return CB_OK;
}

extern "C" int cb_verlet_get_phase_times( const cb_verlet* v, double* ms )
{
    if ( !v || !ms )
        return fail( CB_ERR_INVALID, "cb_verlet_get_phase_times: null argument" );
    const int last = v->prof_slot;
    if ( !v->profiling || !v->have_events || !v->ev_valid[last][0] || !v->ev_valid[last][5] )
        return fail( CB_ERR_INVALID, "cb_verlet_get_phase_times: no profiled build" );
    CB_CUDA( cudaEventSynchronize( v->ev[last][5] ) );
    // average over the (up to kProfRing) most recent builds since profiling was switched on
    double sum[CB_VERLET_NUM_PHASES] = { 0.0 };
    int sets = 0;
    for ( int r = 0; r < cb_verlet::kProfRing; ++r )
    {
        if ( !v->ev_valid[r][0] || !v->ev_valid[r][5] )
            continue;
        for ( int i = 0; i < 5; ++i )
        {
            float t = 0.f;
            if ( v->ev_valid[r][i] && v->ev_valid[r][i + 1] )
                CB_CUDA( cudaEventElapsedTime( &t, v->ev[r][i], v->ev[r][i + 1] ) );
            sum[i] += t;
        }
        float t = 0.f;
        CB_CUDA( cudaEventElapsedTime( &t, v->ev[r][0], v->ev[r][5] ) );
        sum[5] += t;
        ++sets;
    }
    for ( int i = 0; i < CB_VERLET_NUM_PHASES; ++i )
        ms[i] = sum[i] / (double)sets;
    return CB_OK;
}

extern "C" int cb_verlet_get( const cb_verlet* v, cb_verlet_view* view )
{
    if ( !v || !view )
        return fail( CB_ERR_INVALID, "cb_verlet_get: null argument" );
    if ( !v->built )
        return fail( CB_ERR_INVALID, "cb_verlet_get: list not built" );
    view->layout = v->layout;
    view->algorithm = v->algorithm;
    view->n = v->n;
    view->counts = v->counts.as<int32_t>();
    view->offsets = v->layout == CB_LAYOUT_CSR ? v->offsets.as<int32_t>() : nullptr;
    view->neighbors = v->neighbors.as<int32_t>();
    view->total = v->total;
    view->max_n = v->max_n;
    view->width = v->layout == CB_LAYOUT_2D ? v->width : 0;
    view->row_stride = v->layout == CB_LAYOUT_2D ? v->width : 0;
    view->col_stride = v->layout == CB_LAYOUT_2D ? 1 : 0;
    view->refilled = v->refilled;
    view->extent = v->layout == CB_LAYOUT_CSR ? v->extent : 0;
    return CB_OK;
}

extern "C" int cb_verlet_set_row_placement( cb_verlet* v, int placement )
{
    if ( !v || ( placement != CB_ROWS_REFERENCE && placement != CB_ROWS_BINNED ) )
        return fail( CB_ERR_INVALID, "cb_verlet_set_row_placement: bad argument" );
    v->row_placement = placement;
    return CB_OK;
}

extern "C" int cb_verlet_set_neighbor( cb_verlet* v, int64_t i, int64_t k,
                                       int32_t new_index, cb_stream_t stream_ )
{
    if ( !v || !v->built )
        return fail( CB_ERR_INVALID, "cb_verlet_set_neighbor: list not built" );
    if ( i < 0 || i >= v->n || k < 0 )
        return fail( CB_ERR_INVALID, "cb_verlet_set_neighbor: index out of range" );
    k_set_neighbor<<<1, 1, 0, (cudaStream_t)stream_>>>(
        v->neighbors.as<int>(),
        v->layout == CB_LAYOUT_CSR ? v->offsets.as<int>() : nullptr, i, k, v->width,
        new_index );
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_verlet_build_host( cb_verlet* v, const cb_positions* x_h,
                                     int64_t begin, int64_t end, double radius,
                                     double cell_size_ratio, const double* grid_min,
                                     const double* grid_max, int64_t max_neigh,
                                     int algorithm, int layout, int build_op,
                                     cb_stream_t stream_ )
{
    if ( !v || !x_h )
        return fail( CB_ERR_INVALID, "cb_verlet_build_host: null argument" );
    if ( x_h->vlen < 1 || x_h->n < 0 )
        return fail( CB_ERR_INVALID, "cb_verlet_build_host: bad descriptor" );
    cudaStream_t stream = (cudaStream_t)stream_;
    // Number of doubles the descriptor spans.
    size_t span = 1;
    if ( x_h->n > 0 )
    {
        PosAccess a = make_access( *x_h );
        long long last = a.offset( x_h->n - 1 );
        if ( x_h->vlen > 1 ) // whole last SoA
            last = a.outer_stride * ( ( x_h->n - 1 ) / x_h->vlen ) + ( x_h->vlen - 1 );
        span = (size_t)( last + 2 * x_h->comp_stride + 1 );
    }
    CB_TRY( v->host_stage.ensure( span * sizeof( double ), 1.1 ) );
    CB_CUDA( cudaMemcpyAsync( v->host_stage.ptr, x_h->base, span * sizeof( double ),
                              cudaMemcpyHostToDevice, stream ) );
    cb_positions xd = *x_h;
    xd.base = v->host_stage.as<double>();
    return cb_verlet_build( v, &xd, begin, end, radius, cell_size_ratio, grid_min,
                            grid_max, max_neigh, algorithm, layout, build_op, stream_ );
}

extern "C" int cb_verlet_build_2d( cb_verlet* v, const cb_positions* x2, int64_t begin,
                                   int64_t end, double radius, double cell_size_ratio,
                                   const double* grid_min2, const double* grid_max2,
                                   int64_t max_neigh, int algorithm, int layout, int build_op,
                                   cb_stream_t stream_ )
{
    if ( !v || !x2 || !grid_min2 || !grid_max2 )
        return fail( CB_ERR_INVALID, "cb_verlet_build_2d: null argument" );
    if ( x2->vlen < 1 || x2->n < 0 || !( radius > 0.0 ) || !( cell_size_ratio > 0.0 ) )
        return fail( CB_ERR_INVALID, "cb_verlet_build_2d: bad argument" );
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long n = x2->n;
    CB_TRY( v->host_stage.ensure( sizeof( double ) * 3 * (size_t)( n > 0 ? n : 1 ), 1.1 ) );
    if ( n > 0 )
    {
        k_embed_2d<<<launch_grid_for( n, kBlock ), kBlock, 0, stream>>>(
            make_access( *x2 ), n, v->host_stage.as<double>() );
        CB_CHECK_LAUNCH();
    }
    // Third dimension: every particle at z = 0 in the MIDDLE of the middle cell of a z range
    // 3.5 cells wide (floor( 3.5 ) = 3 cells whatever the rounding of delta).  All particles share
    // that z cell; its minDistanceToPoint term and every (dz * dz) are exactly +0.0, so the cell
    // prune, the distances and the x-then-y half criterion are bit for bit the 2-D ones.
    const double delta = cell_size_ratio * radius;
    const double gmin[3] = { grid_min2[0], grid_min2[1], -1.75 * delta };
    const double gmax[3] = { grid_max2[0], grid_max2[1], 1.75 * delta };
    cb_positions x3;
    x3.base = v->host_stage.as<double>();
    x3.n = n;
    x3.outer_stride = 3;
    x3.vlen = 1;
    x3.comp_stride = 1;
    return cb_verlet_build( v, &x3, begin, end, radius, cell_size_ratio, gmin, gmax, max_neigh,
                            algorithm, layout, build_op, stream_ );
}

extern "C" int cb_verlet_copy_to_host( const cb_verlet* v, int32_t* counts_h,
                                       int32_t* offsets_h, int32_t* neighbors_h,
                                       int64_t neighbors_capacity, cb_stream_t stream_ )
{
    if ( !v || !v->built )
        return fail( CB_ERR_INVALID, "cb_verlet_copy_to_host: list not built" );
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long nn =
        v->layout == CB_LAYOUT_CSR ? v->extent : v->n * v->width;
    if ( neighbors_h && neighbors_capacity < nn )
        return fail( CB_ERR_NOMEM, "cb_verlet_copy_to_host: neighbors_h too small" );
    if ( counts_h && v->n > 0 )
        CB_CUDA( cudaMemcpyAsync( counts_h, v->counts.ptr, sizeof( int ) * (size_t)v->n,
                                  cudaMemcpyDeviceToHost, stream ) );
    if ( offsets_h && v->layout == CB_LAYOUT_CSR && v->n > 0 )
        CB_CUDA( cudaMemcpyAsync( offsets_h, v->offsets.ptr,
                                  sizeof( int ) * (size_t)v->n, cudaMemcpyDeviceToHost,
                                  stream ) );
    if ( neighbors_h && nn > 0 )
        CB_CUDA( cudaMemcpyAsync( neighbors_h, v->neighbors.ptr,
                                  sizeof( int ) * (size_t)nn, cudaMemcpyDeviceToHost,
                                  stream ) );
    CB_CUDA( cudaStreamSynchronize( stream ) );
    return CB_OK;
}
