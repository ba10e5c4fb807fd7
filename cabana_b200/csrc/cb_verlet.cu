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

#include "cb_common.cuh"
#include "cb_internal.h"

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
                     double* __restrict__ zs )
{
    for ( long long s = (long long)blockIdx.x * kBlock + threadIdx.x; s < n;
          s += (long long)gridDim.x * kBlock )
    {
        const long long off = x.offset( (long long)permute[s] );
        xs[s] = x.base[off];
        ys[s] = x.base[off + x.comp_stride];
        zs[s] = x.base[off + 2 * x.comp_stride];
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
    int refilled = 0;
    bool built = false;
    DeviceBuffer counts, offsets, neighbors;
    // workspace
    DeviceBuffer cell_counts, cell_off, permute, cell_of, rank, scan, xs, ys, zs, stats;
    DeviceBuffer host_stage; // device copy of host positions (build_host)
    PinnedScalars pinned;
    // optional phase timing
    bool profiling = false;
    cudaEvent_t ev[CB_VERLET_NUM_PHASES + 1] = {};
    bool ev_valid[CB_VERLET_NUM_PHASES + 1] = {};
    bool have_events = false;
    ~cb_verlet()
    {
        if ( have_events )
            for ( auto& e : ev )
                cudaEventDestroy( e );
    }
    // mark(i): boundary i on the stream -- 0 start, 1 after binning, 2 after gather,
    // 3 after count, 4 after scan, 5 after fill.
    void mark( int i, cudaStream_t s )
    {
        if ( profiling && have_events )
        {
            cudaEventRecord( ev[i], s );
            ev_valid[i] = true;
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

extern "C" int cb_verlet_build( cb_verlet* v, const cb_positions* x, int64_t begin,
                                int64_t end, double radius, double cell_size_ratio,
                                const double* grid_min, const double* grid_max,
                                int64_t max_neigh, int algorithm, int layout,
                                int build_op, cb_stream_t stream_ )
{
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

    CB_TRY( v->pinned.ensure() );
    CB_TRY( v->stats.ensure( 4 * sizeof( long long ) ) );
    CB_TRY( v->counts.ensure( sizeof( int ) * na, 1.1 ) );
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
    v->total = 0;
    v->max_n = 0;
    v->width = 0;
    v->refilled = 0;

    for ( auto& f : v->ev_valid )
        f = false;
    v->mark( 0, stream );

    // counts zero-initialised (:215-216)
    CB_CUDA( cudaMemsetAsync( v->counts.ptr, 0, sizeof( int ) * na, stream ) );

    // LCL over ALL particles, not just [begin,end) (:229-235).
    CB_TRY( bin_particles( grid, *x, 0, n, v->cell_counts.as<int>(),
                           v->cell_off.as<unsigned>(), v->permute.as<unsigned>(),
                           v->cell_of.as<int>(), v->rank, v->scan, stream ) );
    v->mark( 1, stream );
    if ( n > 0 )
    {
        k_gather_sorted<<<launch_grid_for( n, kBlock ), kBlock, 0, stream>>>(
            make_access( *x ), n, v->permute.as<unsigned>(), v->xs.as<double>(),
            v->ys.as<double>(), v->zs.as<double>() );
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

    long long* stats_dev = v->stats.as<long long>();
    long long* stats_h = v->pinned.ptr;

    if ( layout == CB_LAYOUT_CSR )
    {
        // count -> processCounts(CSR) -> fill  (:1454-1480, :507-532)
        CB_TRY( launch_pass<kCount>( a, algorithm, layout, stream ) );
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
        if ( v->total > 2147483647ll )
            return fail( CB_ERR_OVERFLOW,
                         "cb_verlet_build: total neighbours exceed INT_MAX" );
        CB_TRY( v->neighbors.ensure( sizeof( int ) * (size_t)( v->total > 0 ? v->total : 1 ),
                                     1.05 ) );
        a.offsets = v->offsets.as<int>();
        a.neighbors = v->neighbors.as<int>();
        CB_TRY( launch_pass<kFill>( a, algorithm, layout, stream ) );
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
            CB_TRY( launch_pass<kFill>( a, algorithm, layout, stream ) );
        }
        else
        {
            CB_TRY( launch_pass<kCount>( a, algorithm, layout, stream ) );
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
        if ( count || v->max_n > v->width )
        {
            v->width = v->max_n;
            if ( (double)na * (double)v->width > 9.0e18 )
                return fail( CB_ERR_OVERFLOW, "cb_verlet_build: 2D list too large" );
            CB_TRY( v->neighbors.ensure(
                sizeof( int ) * na * (size_t)( v->width > 0 ? v->width : 1 ) ) );
            a.neighbors = v->neighbors.as<int>();
            a.width = v->width;
            CB_TRY( launch_pass<kFill>( a, algorithm, layout, stream ) );
            if ( !count )
                v->refilled = 1;
        }
        v->mark( 5, stream );
    }
    v->built = true;
    return CB_OK;
}

extern "C" int cb_verlet_set_profiling( cb_verlet* v, int enable )
{
    if ( !v )
        return fail( CB_ERR_INVALID, "cb_verlet_set_profiling: null argument" );
    if ( enable && !v->have_events )
    {
        for ( auto& e : v->ev )
            CB_CUDA( cudaEventCreate( &e ) );
        v->have_events = true;
    }
    v->profiling = enable != 0;
    return CB_OK;
}

extern "C" int cb_verlet_get_phase_times( const cb_verlet* v, double* ms )
{
    if ( !v || !ms )
        return fail( CB_ERR_INVALID, "cb_verlet_get_phase_times: null argument" );
    if ( !v->profiling || !v->have_events || !v->ev_valid[0] || !v->ev_valid[5] )
        return fail( CB_ERR_INVALID, "cb_verlet_get_phase_times: no profiled build" );
    CB_CUDA( cudaEventSynchronize( v->ev[5] ) );
    for ( int i = 0; i < 5; ++i )
    {
        float t = 0.f;
        if ( v->ev_valid[i] && v->ev_valid[i + 1] )
            CB_CUDA( cudaEventElapsedTime( &t, v->ev[i], v->ev[i + 1] ) );
        ms[i] = t;
    }
    float t = 0.f;
    CB_CUDA( cudaEventElapsedTime( &t, v->ev[0], v->ev[5] ) );
    ms[5] = t;
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

extern "C" int cb_verlet_copy_to_host( const cb_verlet* v, int32_t* counts_h,
                                       int32_t* offsets_h, int32_t* neighbors_h,
                                       int64_t neighbors_capacity, cb_stream_t stream_ )
{
    if ( !v || !v->built )
        return fail( CB_ERR_INVALID, "cb_verlet_copy_to_host: list not built" );
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long nn =
        v->layout == CB_LAYOUT_CSR ? v->total : v->n * v->width;
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
