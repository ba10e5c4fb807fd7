// =============================================================================
// cabana_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain C++/OpenMP restatement of ECP-copa/Cabana's particle neighbour-list
// hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library; the product path
// (cabana_b200/, include/) never links, imports or calls it.
//
// Parity status: PINNED.  The reference itself cannot be compiled in this
// image (every header includes <Kokkos_Core.hpp>; Kokkos is not installed and
// there is no network), so this file restates the algorithm line by line and
// tests/test_oracle_known_answers.py replays every literal known answer the
// reference's own unit tests hold for the path.
//
// Arithmetic contract: IEEE-754 binary64,
// round-to-nearest-even, NO contraction: compile with -ffp-contract=off.
//
// All file:line citations are relative to /root/reference/.
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

// -----------------------------------------------------------------------------
// Position accessor: Cabana slice / rank-2 view layout.
//   element(i,d) = base[outer_stride*(i / vlen) + (i % vlen) + comp_stride*d]
// Slice: outer_stride = Stride, comp_stride = VectorLength
//   (core/src/Cabana_Slice.hpp:134-140, impl/Cabana_Index.hpp:62-80).
// View(n,3): vlen = 1, outer_stride = stride(0), comp_stride = stride(1).
// -----------------------------------------------------------------------------
typedef struct
{
    const double* base;
    int64_t n;
    int64_t outer_stride;
    int32_t vlen;
    int64_t comp_stride;
} orc_positions;

static inline double pos_at( const orc_positions* x, int64_t i, int d )
{
    // vlen is a power of two for every Cabana AoSoA (and 1 for views): shift/mask instead of
    // a 64-bit divide per access (the reference's Impl::Index does the same at compile time,
    // impl/Cabana_Index.hpp:62-80).
    const int64_t v = x->vlen;
    if ( ( v & ( v - 1 ) ) == 0 )
        return x->base[x->outer_stride * ( i >> __builtin_ctzll( (unsigned long long)v ) ) +
                       ( i & ( v - 1 ) ) + x->comp_stride * d];
    return x->base[x->outer_stride * ( i / v ) + ( i % v ) + x->comp_stride * d];
}

// -----------------------------------------------------------------------------
// CartesianGrid  (core/src/impl/Cabana_CartesianGrid.hpp)
// -----------------------------------------------------------------------------
typedef struct
{
    double min[3];
    double max[3];
    double dx[3];
    double rdx[3];
    int nx[3];
} orc_grid;

// cellsBetween: impl/Cabana_CartesianGrid.hpp:279-283
static inline int cells_between( double max, double min, double rdelta )
{
    return (int)std::floor( ( max - min ) * rdelta );
}

// init: impl/Cabana_CartesianGrid.hpp:61-73
void orc_grid_init( orc_grid* g, const double* min, const double* max,
                    const double* delta )
{
    for ( int d = 0; d < 3; ++d )
    {
        g->min[d] = min[d];
        g->max[d] = max[d];
        g->nx[d] = cells_between( max[d], min[d], 1.0 / delta[d] );
        g->dx[d] = ( max[d] - min[d] ) / g->nx[d];
        g->rdx[d] = 1.0 / g->dx[d];
    }
}

// locatePoint: impl/Cabana_CartesianGrid.hpp:171-182
void orc_grid_locate( const orc_grid* g, const double* p, int* c )
{
    for ( int d = 0; d < 3; ++d )
    {
        c[d] = cells_between( p[d], g->min[d], g->rdx[d] );
        c[d] = ( c[d] == g->nx[d] ) ? c[d] - 1 : c[d];
    }
}

// minDistanceToPoint: impl/Cabana_CartesianGrid.hpp:207-223
double orc_grid_min_distance( const orc_grid* g, const double* x,
                              const int* c )
{
    double rsqr = 0.0;
    for ( int d = 0; d < 3; ++d )
    {
        double xc = g->min[d] + ( c[d] + 0.5 ) * g->dx[d];
        double rx = std::fabs( x[d] - xc ) - 0.5 * g->dx[d];
        rx = ( rx > 0.0 ) ? rx : 0.0;
        rsqr += rx * rx;
    }
    return rsqr;
}

// cardinalCellIndex / ijkBinIndex: impl/Cabana_CartesianGrid.hpp:226-276
static inline int cardinal( const orc_grid* g, int i, int j, int k )
{
    return ( i * g->nx[1] + j ) * g->nx[2] + k;
}
int orc_grid_cardinal( const orc_grid* g, int i, int j, int k )
{
    return cardinal( g, i, j, k );
}
void orc_grid_ijk( const orc_grid* g, int c, int* ijk )
{
    ijk[0] = c / ( g->nx[1] * g->nx[2] );
    ijk[1] = ( c / g->nx[2] ) % g->nx[1];
    ijk[2] = c % g->nx[2];
}

// -----------------------------------------------------------------------------
// LinkedCellStencil  (core/src/Cabana_LinkedCellList.hpp:55-66, :105-119)
// -----------------------------------------------------------------------------
typedef struct
{
    orc_grid grid;
    int max_cells_dir;
    int max_cells;
    int cell_range;
} orc_stencil;

void orc_stencil_init( orc_stencil* s, double radius, double ratio,
                       const double* min, const double* max )
{
    double dx = radius * ratio; // :60
    double delta[3] = { dx, dx, dx };
    orc_grid_init( &s->grid, min, max, delta );
    s->cell_range = (int)std::ceil( 1 / ratio ); // :63
    s->max_cells_dir = 2 * s->cell_range + 1;
    s->max_cells = s->max_cells_dir * s->max_cells_dir * s->max_cells_dir;
}

// getCells: :105-119
void orc_stencil_cells( const orc_stencil* s, int cell, int* mn, int* mx )
{
    int ijk[3];
    orc_grid_ijk( &s->grid, cell, ijk );
    for ( int d = 0; d < 3; ++d )
    {
        mn[d] = ( ijk[d] - s->cell_range > 0 ) ? ijk[d] - s->cell_range : 0;
        mx[d] = ( ijk[d] + s->cell_range + 1 < s->grid.nx[d] )
                    ? ijk[d] + s->cell_range + 1
                    : s->grid.nx[d];
    }
}

// -----------------------------------------------------------------------------
// LinkedCellList::build  (core/src/Cabana_LinkedCellList.hpp:651-739) and
// storeParticleBins (:778-783, :806-823).
//
// counts[ncell] (int), offsets[ncell], permute[end-begin] (absolute ids),
// particle_bins[end-begin] (indexed relative to begin, unsorted convention).
// Within-cell order is unspecified in the reference (atomic slot claim, :726);
// here it is ascending particle id when `parallel` == 0 (what a serial backend
// produces) and arbitrary when `parallel` != 0 (OpenMP atomics).
// -----------------------------------------------------------------------------
} // extern "C"

// Exclusive prefix sum (Kokkos::parallel_scan in the reference); chunked two-level scan when
// `parallel`.  Returns the total.
template <class Out>
static int64_t exclusive_scan_i64( const int* in, Out* out, int64_t n, int parallel )
{
    int nt = 1;
#ifdef _OPENMP
    if ( parallel )
        nt = omp_get_max_threads();
#endif
    if ( nt <= 1 || n < 65536 )
    {
        int64_t update = 0;
        for ( int64_t c = 0; c < n; ++c )
        {
            out[c] = (Out)update;
            update += in[c];
        }
        return update;
    }
    std::vector<int64_t> part( nt + 1, 0 );
    const int64_t chunk = ( n + nt - 1 ) / nt;
#pragma omp parallel num_threads( nt )
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        const int64_t lo = std::min<int64_t>( n, chunk * t );
        const int64_t hi = std::min<int64_t>( n, lo + chunk );
        int64_t sum = 0;
        for ( int64_t c = lo; c < hi; ++c )
            sum += in[c];
        part[t + 1] = sum;
#pragma omp barrier
#pragma omp single
        for ( int k = 0; k < nt; ++k )
            part[k + 1] += part[k];
        int64_t update = part[t];
        for ( int64_t c = lo; c < hi; ++c )
        {
            out[c] = (Out)update;
            update += in[c];
        }
    }
    return part[nt];
}

extern "C" {

void orc_lcl_build( const orc_grid* g, const orc_positions* x, int64_t begin,
                    int64_t end, int* counts, int64_t* offsets,
                    int64_t* permute, int* particle_bins, int parallel )
{
    const int64_t ncell = (int64_t)g->nx[0] * g->nx[1] * g->nx[2];
    std::fill( counts, counts + ncell, 0 );

    const int64_t np = end - begin;
    std::vector<int> cell_of( np );

    // cell_count (:683-699)
#pragma omp parallel for schedule( static ) if ( parallel )
    for ( int64_t p = begin; p < end; ++p )
    {
        double pos[3];
        int ijk[3];
        for ( int d = 0; d < 3; ++d )
            pos[d] = pos_at( x, p, d );
        orc_grid_locate( g, pos, ijk );
        int c = cardinal( g, ijk[0], ijk[1], ijk[2] );
        cell_of[p - begin] = c;
#pragma omp atomic
        counts[c] += 1;
    }

    // offset_scan (:701-712): exclusive prefix sum.
    exclusive_scan_i64( counts, offsets, ncell, parallel );

    // reset + create_permute (:715-731)
    std::fill( counts, counts + ncell, 0 );
    if ( parallel )
    {
#pragma omp parallel for schedule( static )
        for ( int64_t p = begin; p < end; ++p )
        {
            int cid = cell_of[p - begin];
            int c;
#pragma omp atomic capture
            c = counts[cid]++;
            permute[offsets[cid] + c] = p;
        }
    }
    else
    {
        for ( int64_t p = begin; p < end; ++p )
        {
            int cid = cell_of[p - begin];
            int c = counts[cid]++;
            permute[offsets[cid] + c] = p;
        }
    }

    // storeParticleBins, unsorted (:806-823)
    if ( particle_bins )
    {
#pragma omp parallel for schedule( static ) if ( parallel )
        for ( int64_t c = 0; c < ncell; ++c )
            for ( int64_t p = offsets[c]; p < offsets[c] + counts[c]; ++p )
                particle_bins[permute[p] - begin] = (int)c;
    }
}

// permute(BinningData, slice)  (core/src/Cabana_Sort.hpp:600-656):
//   tmp[i-begin] = data[perm[i-begin]];  data[i] = tmp[i-begin]  for i in
//   [begin,end).  `data` is a slice-laid-out double field with num_comp
//   components.
void orc_permute_slice( double* data, int64_t outer_stride, int32_t vlen,
                        int64_t comp_stride, int num_comp, int64_t begin,
                        int64_t end, const int64_t* permute )
{
    const int64_t np = end - begin;
    std::vector<double> scratch( (size_t)np * num_comp );
    for ( int64_t i = begin; i < end; ++i )
    {
        int64_t pi = permute[i - begin];
        int64_t off = outer_stride * ( pi / vlen ) + ( pi % vlen );
        for ( int n = 0; n < num_comp; ++n )
            scratch[( i - begin ) * num_comp + n] = data[off + comp_stride * n];
    }
    for ( int64_t i = begin; i < end; ++i )
    {
        int64_t off = outer_stride * ( i / vlen ) + ( i % vlen );
        for ( int n = 0; n < num_comp; ++n )
            data[off + comp_stride * n] = scratch[( i - begin ) * num_comp + n];
    }
}

// -----------------------------------------------------------------------------
// NeighborDiscriminator  (core/src/Cabana_NeighborList.hpp:86-214)
// -----------------------------------------------------------------------------
enum
{
    ORC_FULL = 0,
    ORC_HALF = 1
};
enum
{
    ORC_CSR = 0,
    ORC_2D = 1
};

static inline bool is_valid( int algo, int64_t p, const double* xp, int64_t n,
                             const double* xn )
{
    if ( algo == ORC_FULL )
        return p != n; // :98-104
    // Half: :139-149
    return ( ( p != n ) &&
             ( ( xn[0] > xp[0] ) ||
               ( ( xn[0] == xp[0] ) &&
                 ( ( xn[1] > xp[1] ) ||
                   ( ( xn[1] == xp[1] ) && ( xn[2] > xp[2] ) ) ) ) ) );
}

// -----------------------------------------------------------------------------
// VerletList build  (core/src/Cabana_VerletList.hpp:205-240 init, :316-474
// count, :493-562 processCounts, :572-713 fill, :1441-1484 buildImpl).
//
// Output ownership: the oracle mallocs *neighbors_out (free with orc_free).
// CSR : counts[n], offsets[n], neighbors[total]
// 2D  : counts[n], neighbors[n][width] ROW-MAJOR (host Kokkos default layout),
//       *width_out = allocated extent(1), *max_n_out = _data.max_n.
// -----------------------------------------------------------------------------
typedef struct
{
    int64_t total;   // CSR: neighbors.extent(0); 2D: sum(counts)
    int64_t max_n;   // max(counts)
    int64_t width;   // 2D allocated row width (extent(1)); CSR: 0
    int32_t refilled; // 2D: 1 when the realloc+refill path ran
} orc_verlet_info;

} // extern "C" (templates below need C++ linkage)

struct Builder
{
    const orc_positions* x;
    int64_t begin, end;
    int algo, layout;
    double rsqr;
    orc_grid grid;
    orc_stencil stencil;
    std::vector<int> lcl_counts;
    std::vector<int64_t> lcl_offsets;
    std::vector<int64_t> lcl_permute;
    int* counts;
    int* offsets;
    int* neighbors;
    int64_t width; // 2D extent(1)
};

// neighbor_kernel (count :445-474 / fill :684-713) over one stencil cell.
template <bool Fill>
static inline void
process_bin( Builder& b, int64_t pid, const double* xp, const int* ijk,
             int& stencil_count )
{
    // linkedcell_bin (:394-415 / :641-657): cell-level prune.
    if ( !( orc_grid_min_distance( &b.stencil.grid, xp, ijk ) <= b.rsqr ) )
        return;
    int cell = cardinal( &b.grid, ijk[0], ijk[1], ijk[2] );
    int64_t n_offset = b.lcl_offsets[cell];
    int num_n = b.lcl_counts[cell];
    for ( int n = 0; n < num_n; ++n )
    {
        int64_t nid = b.lcl_permute[n_offset + n];
        double xn[3];
        for ( int d = 0; d < 3; ++d )
            xn[d] = pos_at( b.x, nid, d );
        if ( is_valid( b.algo, pid, xp, nid, xn ) )
        {
            double dist_sqr = 0.0;
            for ( int d = 0; d < 3; ++d )
            {
                double dx = xp[d] - xn[d];
                dist_sqr += dx * dx;
            }
            if ( dist_sqr <= b.rsqr ) // withinCutoff :244-255
            {
                if ( !Fill )
                    stencil_count += 1;
                else
                {
                    // addNeighbor :66-71 (CSR) / :99-105 (2D).  Only row pid is
                    // touched for a fixed radius, and exactly one thread owns
                    // pid, so no atomic is needed here.
                    int count = b.counts[pid]++;
                    if ( b.layout == ORC_CSR )
                        b.neighbors[b.offsets[pid] + count] = (int)nid;
                    else if ( (int64_t)count < b.width )
                        b.neighbors[pid * b.width + count] = (int)nid;
                }
            }
        }
    }
}

template <bool Fill>
static void run_pass( Builder& b )
{
    const int ncell = b.grid.nx[0] * b.grid.nx[1] * b.grid.nx[2];
    // TeamPolicy league = cells, Schedule<Dynamic> (:311-314, :1452-1459)
#pragma omp parallel for schedule( dynamic, 16 )
    for ( int cell = 0; cell < ncell; ++cell )
    {
        int mn[3], mx[3];
        orc_stencil_cells( &b.stencil, cell, mn, mx );
        int64_t b_offset = b.lcl_offsets[cell];
        for ( int bi = 0; bi < b.lcl_counts[cell]; ++bi )
        {
            int64_t pid = b.lcl_permute[bi + b_offset];
            if ( ( pid >= b.begin ) && ( pid < b.end ) )
            {
                double xp[3];
                for ( int d = 0; d < 3; ++d )
                    xp[d] = pos_at( b.x, pid, d );
                int stencil_count = 0;
                int ijk[3];
                for ( int i = mn[0]; i < mx[0]; ++i )
                    for ( int j = mn[1]; j < mx[1]; ++j )
                        for ( int k = mn[2]; k < mx[2]; ++k )
                        {
                            ijk[0] = i;
                            ijk[1] = j;
                            ijk[2] = k;
                            process_bin<Fill>( b, pid, xp, ijk, stencil_count );
                        }
                if ( !Fill )
                    b.counts[pid] = stencil_count;
            }
        }
    }
}

extern "C" {

int orc_verlet_build( const orc_positions* x, int64_t begin, int64_t end,
                      double radius, double cell_size_ratio,
                      const double* grid_min, const double* grid_max,
                      int64_t max_neigh, int algo, int layout, int* counts,
                      int* offsets, int** neighbors_out,
                      orc_verlet_info* info )
{
    Builder b;
    b.x = x;
    b.begin = begin;
    b.end = end;
    b.algo = algo;
    b.layout = layout;
    b.counts = counts;
    b.offsets = offsets;
    b.neighbors = nullptr;
    b.width = 0;
    const int64_t n = x->n;

    // init (:205-240)
    bool count = true;
    bool refill = false;
    std::fill( counts, counts + n, 0 );
    if ( layout == ORC_2D && max_neigh > 0 ) // initCounts(2D) :495-505
    {
        count = false;
        b.width = max_neigh;
        b.neighbors = (int*)std::malloc( sizeof( int ) * (size_t)n *
                                         (size_t)b.width );
    }
    double grid_size = cell_size_ratio * radius; // :224
    double delta[3] = { grid_size, grid_size, grid_size };
    orc_grid_init( &b.grid, grid_min, grid_max, delta );
    orc_stencil_init( &b.stencil, radius, cell_size_ratio, grid_min,
                      grid_max );
    const int64_t ncell = (int64_t)b.grid.nx[0] * b.grid.nx[1] * b.grid.nx[2];
    b.lcl_counts.resize( ncell );
    b.lcl_offsets.resize( ncell );
    b.lcl_permute.resize( n );
    // LCL bins ALL particles (:229-235).
    // (threaded like the reference's OpenMP backend whenever more than one thread is
    // available: within-cell order is unspecified there, :726)
    int par = 0;
#ifdef _OPENMP
    par = omp_get_max_threads() > 1 ? 1 : 0;
#endif
    orc_lcl_build( &b.grid, x, 0, n, b.lcl_counts.data(), b.lcl_offsets.data(),
                   b.lcl_permute.data(), nullptr, par );
    b.rsqr = radius * radius; // :239

    info->refilled = 0;

    // buildImpl (:1441-1484)
    if ( count )
        run_pass<false>( b );
    else
        run_pass<true>( b ); // processCounts(2D) before it is a no-op (:1463)

    // processCounts
    if ( layout == ORC_CSR ) // :507-532
    {
        const int64_t total = exclusive_scan_i64( counts, offsets, n, par );
        b.neighbors =
            (int*)std::malloc( sizeof( int ) * (size_t)std::max<int64_t>( total, 1 ) );
        std::fill( counts, counts + n, 0 );
        info->total = total;
    }
    else // :536-562
    {
        int mx = 0;
        for ( int64_t i = 0; i < n; ++i )
            if ( counts[i] > mx )
                mx = counts[i];
        info->max_n = mx;
        if ( count || (int64_t)mx > b.width )
        {
            refill = true;
            std::fill( counts, counts + n, 0 );
            if ( b.neighbors )
                std::free( b.neighbors );
            b.width = mx;
            b.neighbors = (int*)std::malloc(
                sizeof( int ) * (size_t)std::max<int64_t>( n * b.width, 1 ) );
        }
    }

    if ( count || refill )
    {
        run_pass<true>( b );
        if ( !count )
            info->refilled = 1;
    }

    int64_t mxn = 0, tot = 0;
    for ( int64_t i = 0; i < n; ++i )
    {
        if ( counts[i] > mxn )
            mxn = counts[i];
        tot += counts[i];
    }
    info->max_n = mxn;
    if ( layout == ORC_2D )
        info->total = tot;
    info->width = b.width;
    *neighbors_out = b.neighbors;
    return 0;
}

// -----------------------------------------------------------------------------
// VerletList build with a PER-PARTICLE cutoff radius
// (core/src/Cabana_VerletList.hpp:181-203 builder ctor; :244-305 withinCutoff /
//  neighborNotWithinCutoff / countNeighbor / addNeighbor; :989-1017 list ctor).
//
// The LinkedCellList and the stencil use `background_radius`; the cell prune and
// the pair test use radius(i) of the particle whose row is being built (:401-403,
// :254).  A pair found by i with dist_sqr >= radius(j)^2 "will not be found by j",
// so i also appears in j's row (:299-304).
//
// Reference quirk: the COUNT pass adds that extra entry
// to i's count (:277-290) while the FILL pass appends it to j's row (:293-305).
// The CSR offsets are scanned from the count-pass values, so rows can overrun
// their slots; only the totals agree.  The counts after the fill pass -- which is
// what the reference's own test checks (tstNeighborList.hpp:244-252) -- are the
// row sizes of the fill semantics.  This restatement returns
//   counts / offsets / neighbors : the FILL semantics (rows as addNeighbor builds
//                                  them, offsets = exclusive scan of those sizes),
//   count_pass_counts (optional) : the literal count-pass values (:277-290),
// so a test can show both the 6/4 answer and the asymmetry.
// 2D: width = max(max_neigh, max row size) (no entry is ever dropped).
// -----------------------------------------------------------------------------
int orc_verlet_build_radii( const orc_positions* x, const double* radii, int64_t begin,
                            int64_t end, double background_radius, double cell_size_ratio,
                            const double* grid_min, const double* grid_max,
                            int64_t max_neigh, int algo, int layout, int* counts,
                            int* offsets, int** neighbors_out, int* count_pass_counts,
                            orc_verlet_info* info )
{
    const int64_t n = x->n;
    orc_grid grid;
    orc_stencil stencil;
    double grid_size = cell_size_ratio * background_radius; // :224
    double delta[3] = { grid_size, grid_size, grid_size };
    orc_grid_init( &grid, grid_min, grid_max, delta );
    orc_stencil_init( &stencil, background_radius, cell_size_ratio, grid_min, grid_max );
    const int64_t ncell = (int64_t)grid.nx[0] * grid.nx[1] * grid.nx[2];
    std::vector<int> lcl_counts( ncell );
    std::vector<int64_t> lcl_offsets( ncell ), lcl_permute( n );
    orc_lcl_build( &grid, x, 0, n, lcl_counts.data(), lcl_offsets.data(), lcl_permute.data(),
                   nullptr, 0 );
    std::vector<std::vector<int>> rows( (size_t)n );
    if ( count_pass_counts )
        std::fill( count_pass_counts, count_pass_counts + n, 0 );
    for ( int64_t cell = 0; cell < ncell; ++cell )
    {
        int mn[3], mx[3];
        orc_stencil_cells( &stencil, (int)cell, mn, mx );
        for ( int bi = 0; bi < lcl_counts[cell]; ++bi )
        {
            const int64_t pid = lcl_permute[bi + lcl_offsets[cell]];
            if ( pid < begin || pid >= end )
                continue;
            double xp[3];
            for ( int d = 0; d < 3; ++d )
                xp[d] = pos_at( x, pid, d );
            const double ri2 = radii[pid] * radii[pid]; // withinCutoff :252
            int ijk[3];
            for ( ijk[0] = mn[0]; ijk[0] < mx[0]; ++ijk[0] )
                for ( ijk[1] = mn[1]; ijk[1] < mx[1]; ++ijk[1] )
                    for ( ijk[2] = mn[2]; ijk[2] < mx[2]; ++ijk[2] )
                    {
                        if ( !( orc_grid_min_distance( &stencil.grid, xp, ijk ) <= ri2 ) )
                            continue; // :401-403 with the particle's own radius
                        const int c = cardinal( &grid, ijk[0], ijk[1], ijk[2] );
                        for ( int k = 0; k < lcl_counts[c]; ++k )
                        {
                            const int64_t nid = lcl_permute[lcl_offsets[c] + k];
                            double xn[3];
                            for ( int d = 0; d < 3; ++d )
                                xn[d] = pos_at( x, nid, d );
                            if ( !is_valid( algo, pid, xp, nid, xn ) )
                                continue;
                            double dist_sqr = 0.0;
                            for ( int d = 0; d < 3; ++d )
                            {
                                double dx = xp[d] - xn[d];
                                dist_sqr += dx * dx;
                            }
                            if ( dist_sqr <= ri2 ) // :299
                            {
                                rows[(size_t)pid].push_back( (int)nid ); // :301
                                if ( count_pass_counts )
                                    count_pass_counts[pid] += 1; // :284
                                if ( dist_sqr >= radii[nid] * radii[nid] ) // :268
                                {
                                    rows[(size_t)nid].push_back( (int)pid ); // :304
                                    if ( count_pass_counts )
                                        count_pass_counts[pid] += 1; // :288 (sic: i, not j)
                                }
                            }
                        }
                    }
        }
    }
    int64_t total = 0, mxn = 0;
    for ( int64_t i = 0; i < n; ++i )
    {
        counts[i] = (int)rows[(size_t)i].size();
        if ( layout == ORC_CSR )
            offsets[i] = (int)total;
        total += counts[i];
        mxn = std::max<int64_t>( mxn, counts[i] );
    }
    info->total = total;
    info->max_n = mxn;
    info->refilled = 0;
    info->width = 0;
    int* nb = nullptr;
    if ( layout == ORC_CSR )
    {
        nb = (int*)std::malloc( sizeof( int ) * (size_t)std::max<int64_t>( total, 1 ) );
        for ( int64_t i = 0; i < n; ++i )
            std::copy( rows[(size_t)i].begin(), rows[(size_t)i].end(), nb + offsets[i] );
    }
    else
    {
        int64_t width = mxn;
        if ( max_neigh > 0 && mxn <= max_neigh )
            width = max_neigh;
        else if ( max_neigh > 0 )
            info->refilled = 1;
        info->width = width;
        nb = (int*)std::calloc( (size_t)std::max<int64_t>( n * width, 1 ), sizeof( int ) );
        for ( int64_t i = 0; i < n; ++i )
            std::copy( rows[(size_t)i].begin(), rows[(size_t)i].end(), nb + i * width );
    }
    *neighbors_out = nb;
    return 0;
}

// -----------------------------------------------------------------------------
// 2-D VerletList build (NumSpaceDim = 2: Cabana_VerletList.hpp:377-392 / :626-639 stencil
// loops, impl/Cabana_CartesianGrid.hpp for two dimensions, the 2-D cardinal index
// i * ny + j (:268-275), the 2-D half criterion Cabana_NeighborList.hpp:185-190).
// Positions: (n,2) elements described like every slice; CSR layout only (the 2D-layout
// bookkeeping is dimension independent and covered by orc_verlet_build).
// -----------------------------------------------------------------------------
int orc_verlet_build_2d( const orc_positions* x, int64_t begin, int64_t end, double radius,
                         double cell_size_ratio, const double* grid_min, const double* grid_max,
                         int algo, int* counts, int* offsets, int** neighbors_out,
                         orc_verlet_info* info )
{
    const int64_t n = x->n;
    // CartesianGrid<double,2>::init (:61-73) for the binning grid (delta = ratio * r, :224) and
    // the stencil grid (r * ratio, Cabana_LinkedCellList.hpp:60): the same doubles.
    const double delta = cell_size_ratio * radius;
    int nx[2];
    double dx[2], rdx[2];
    for ( int d = 0; d < 2; ++d )
    {
        nx[d] = cells_between( grid_max[d], grid_min[d], 1.0 / delta );
        dx[d] = ( grid_max[d] - grid_min[d] ) / nx[d];
        rdx[d] = 1.0 / dx[d];
    }
    const int cell_range = (int)std::ceil( 1 / cell_size_ratio );
    auto locate = [&]( double p, int d )
    {
        int c = cells_between( p, grid_min[d], rdx[d] ); // :171-182
        return c == nx[d] ? c - 1 : c;
    };
    auto min_dist = [&]( const double* xp, int i, int j )
    {
        // minDistanceToPoint (:207-223), two dimensions
        const int ij[2] = { i, j };
        double rsqr = 0.0;
        for ( int d = 0; d < 2; ++d )
        {
            const double xc = grid_min[d] + ( ij[d] + 0.5 ) * dx[d];
            const double rx = std::fabs( xp[d] - xc ) - 0.5 * dx[d];
            const double r = rx > 0.0 ? rx : 0.0;
            rsqr += r * r;
        }
        return rsqr;
    };
    const int64_t ncell = (int64_t)nx[0] * nx[1];
    std::vector<int> cc( ncell, 0 );
    std::vector<int64_t> coff( ncell + 1, 0 ), perm( n ), cell_of( n );
    for ( int64_t p = 0; p < n; ++p )
    {
        const int i = locate( pos_at( x, p, 0 ), 0 ), j = locate( pos_at( x, p, 1 ), 1 );
        cell_of[p] = (int64_t)i * nx[1] + j;
        ++cc[cell_of[p]];
    }
    for ( int64_t c = 0; c < ncell; ++c )
        coff[c + 1] = coff[c] + cc[c];
    {
        std::vector<int64_t> fill( coff.begin(), coff.end() - 1 );
        for ( int64_t p = 0; p < n; ++p )
            perm[fill[cell_of[p]]++] = p;
    }
    const double rsqr = radius * radius;
    std::vector<std::vector<int>> rows( (size_t)n );
    for ( int64_t cell = 0; cell < ncell; ++cell )
    {
        const int ci = (int)( cell / nx[1] ), cj = (int)( cell % nx[1] );
        const int imin = ci - cell_range > 0 ? ci - cell_range : 0;
        const int imax = ci + cell_range + 1 < nx[0] ? ci + cell_range + 1 : nx[0];
        const int jmin = cj - cell_range > 0 ? cj - cell_range : 0;
        const int jmax = cj + cell_range + 1 < nx[1] ? cj + cell_range + 1 : nx[1];
        for ( int64_t b = coff[cell]; b < coff[cell + 1]; ++b )
        {
            const int64_t pid = perm[b];
            if ( pid < begin || pid >= end )
                continue;
            const double xp[2] = { pos_at( x, pid, 0 ), pos_at( x, pid, 1 ) };
            for ( int i = imin; i < imax; ++i )
                for ( int j = jmin; j < jmax; ++j )
                {
                    if ( !( min_dist( xp, i, j ) <= rsqr ) )
                        continue;
                    const int64_t c = (int64_t)i * nx[1] + j;
                    for ( int64_t k = coff[c]; k < coff[c + 1]; ++k )
                    {
                        const int64_t nid = perm[k];
                        const double xn[2] = { pos_at( x, nid, 0 ), pos_at( x, nid, 1 ) };
                        bool valid = pid != nid;
                        if ( algo == ORC_HALF ) // :185-190
                            valid = valid && ( ( xn[0] > xp[0] ) ||
                                               ( ( xn[0] == xp[0] ) && ( xn[1] > xp[1] ) ) );
                        if ( !valid )
                            continue;
                        double dist_sqr = 0.0;
                        for ( int d = 0; d < 2; ++d )
                        {
                            const double dd = xp[d] - xn[d];
                            dist_sqr += dd * dd;
                        }
                        if ( dist_sqr <= rsqr )
                            rows[(size_t)pid].push_back( (int)nid );
                    }
                }
        }
    }
    int64_t total = 0, mxn = 0;
    for ( int64_t i = 0; i < n; ++i )
    {
        counts[i] = (int)rows[(size_t)i].size();
        offsets[i] = (int)total;
        total += counts[i];
        mxn = std::max<int64_t>( mxn, counts[i] );
    }
    int* nb = (int*)std::malloc( sizeof( int ) * (size_t)std::max<int64_t>( total, 1 ) );
    for ( int64_t i = 0; i < n; ++i )
        std::copy( rows[(size_t)i].begin(), rows[(size_t)i].end(), nb + offsets[i] );
    info->total = total;
    info->max_n = mxn;
    info->width = 0;
    info->refilled = 0;
    *neighbors_out = nb;
    return 0;
}

void orc_free( void* p ) { std::free( p ); }

// Order-independent 64-bit hash of every row (multiset of neighbour ids): equal hashes for
// all rows <=> equal sorted rows, without sorting 1e9 ids.  Test infrastructure for the
// full-size parity checks.  layout CSR: row i = neighbors[offsets[i] .. +counts[i]);
// 2D: neighbors[i*width .. +counts[i]).
static inline uint64_t mix64( uint64_t z )
{
    z += 0x9e3779b97f4a7c15ull;
    z = ( z ^ ( z >> 30 ) ) * 0xbf58476d1ce4e5b9ull;
    z = ( z ^ ( z >> 27 ) ) * 0x94d049bb133111ebull;
    return z ^ ( z >> 31 );
}
void orc_row_hashes( int layout, int64_t n, const int* counts, const int* offsets,
                     const int* neighbors, int64_t width, uint64_t* out )
{
#pragma omp parallel for schedule( static )
    for ( int64_t i = 0; i < n; ++i )
    {
        const int* row =
            neighbors + ( layout == ORC_CSR ? (int64_t)offsets[i] : i * width );
        uint64_t h = 0x1234567ull * (uint64_t)counts[i];
        for ( int k = 0; k < counts[i]; ++k )
            h += mix64( (uint64_t)(uint32_t)row[k] );
        out[i] = h;
    }
}

// -----------------------------------------------------------------------------
// Brute-force N^2 list  (core/unit_test/neighbor_unit_test.hpp:86-158).
// counts[n]; neighbors row-major [n][width]; pass neighbors == NULL to count.
// -----------------------------------------------------------------------------
void orc_brute_force( const orc_positions* x, double radius, int* counts,
                      int* neighbors, int64_t width )
{
    const int64_t n = x->n;
    const double rsqr = radius * radius;
#pragma omp parallel for schedule( static )
    for ( int64_t i = 0; i < n; ++i )
    {
        int c = 0;
        for ( int64_t j = 0; j < n; ++j )
        {
            if ( i != j )
            {
                double dsqr = 0.0;
                for ( int d = 0; d < 3; ++d )
                    dsqr += ( pos_at( x, i, d ) - pos_at( x, j, d ) ) *
                            ( pos_at( x, i, d ) - pos_at( x, j, d ) );
                if ( dsqr <= rsqr )
                {
                    if ( neighbors && c < width )
                        neighbors[i * width + c] = (int)j;
                    ++c;
                }
            }
        }
        counts[i] = c;
    }
}

// -----------------------------------------------------------------------------
// neighbor_parallel_for with the Lennard-Jones consumer.
// Traversal loop = core/src/Cabana_Parallel.hpp:280-288 (Serial) / :416-430
// (Team): for i in [begin,end): for n in [0,numNeighbor(i)): f(i, nbr(i,n)).
// The LJ functor itself is this repo's (north_star): for pair (i,j) with
// r2 = |x_i - x_j|^2 < rc^2:
//     sr2 = sigma^2 / r2; sr6 = sr2^3
//     fpair = 24 eps sr6 (2 sr6 - 1) / r2
//     f_i += fpair (x_i - x_j)         (and f_j -= ... when `newton` != 0)
// fabs_out (optional) accumulates sum_j |fpair (x_i - x_j)| per component, the
// scale against which the 1e-12 relative tolerance is checked.
// -----------------------------------------------------------------------------
static inline int list_num( int layout, const int* counts, int64_t i )
{
    (void)layout;
    return counts[i];
}
static inline int list_get( int layout, const int* offsets,
                            const int* neighbors, int64_t width, int64_t i,
                            int n )
{
    return layout == ORC_CSR ? neighbors[offsets[i] + n]
                             : neighbors[i * width + n];
}

void orc_lj_forces( const orc_positions* x, int layout, const int* counts,
                    const int* offsets, const int* neighbors, int64_t width,
                    int64_t begin, int64_t end, double eps, double sigma,
                    double rc, int newton, double* f, double* fabs_out )
{
    const double rc2 = rc * rc;
    const double s2 = sigma * sigma;
    // Serial when newton (scatter to j) so the oracle stays deterministic.
#pragma omp parallel for schedule( static ) if ( !newton )
    for ( int64_t i = begin; i < end; ++i )
    {
        double xi[3] = { pos_at( x, i, 0 ), pos_at( x, i, 1 ),
                         pos_at( x, i, 2 ) };
        int nn = list_num( layout, counts, i );
        for ( int n = 0; n < nn; ++n )
        {
            int64_t j = list_get( layout, offsets, neighbors, width, i, n );
            double d[3] = { xi[0] - pos_at( x, j, 0 ),
                            xi[1] - pos_at( x, j, 1 ),
                            xi[2] - pos_at( x, j, 2 ) };
            double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            if ( r2 < rc2 )
            {
                double sr2 = s2 / r2;
                double sr6 = sr2 * sr2 * sr2;
                double fpair = 24.0 * eps * sr6 * ( 2.0 * sr6 - 1.0 ) / r2;
                for ( int c = 0; c < 3; ++c )
                {
                    double fc = fpair * d[c];
                    f[3 * i + c] += fc;
                    if ( fabs_out )
                        fabs_out[3 * i + c] += std::fabs( fc );
                    if ( newton )
                    {
                        f[3 * j + c] -= fc;
                        if ( fabs_out )
                            fabs_out[3 * j + c] += std::fabs( fc );
                    }
                }
            }
        }
    }
}

// neighbor_parallel_reduce (core/src/Cabana_Parallel.hpp:638-685, :787-844)
// with the LJ pair-energy consumer: sum over stored pairs of
// scale * 4 eps (sr6^2 - sr6), r2 < rc^2.  Callers pass scale = 0.5 for full
// lists and 1.0 for half lists.
double orc_lj_energy( const orc_positions* x, int layout, const int* counts,
                      const int* offsets, const int* neighbors, int64_t width,
                      int64_t begin, int64_t end, double eps, double sigma,
                      double rc, double scale )
{
    const double rc2 = rc * rc;
    const double s2 = sigma * sigma;
    double total = 0.0;
#pragma omp parallel for schedule( static ) reduction( + : total )
    for ( int64_t i = begin; i < end; ++i )
    {
        double xi[3] = { pos_at( x, i, 0 ), pos_at( x, i, 1 ),
                         pos_at( x, i, 2 ) };
        int nn = list_num( layout, counts, i );
        double e = 0.0;
        for ( int n = 0; n < nn; ++n )
        {
            int64_t j = list_get( layout, offsets, neighbors, width, i, n );
            double d0 = xi[0] - pos_at( x, j, 0 );
            double d1 = xi[1] - pos_at( x, j, 1 );
            double d2 = xi[2] - pos_at( x, j, 2 );
            double r2 = d0 * d0 + d1 * d1 + d2 * d2;
            if ( r2 < rc2 )
            {
                double sr2 = s2 / r2;
                double sr6 = sr2 * sr2 * sr2;
                e += 4.0 * eps * ( sr6 * sr6 - sr6 );
            }
        }
        total += scale * e;
    }
    return total;
}

// The reference unit tests' traversal functor
// (core/unit_test/neighbor_unit_test.hpp:291-419): result[i] += j for every
// stored neighbour j of i.
void orc_neighbor_id_sum( int layout, const int* counts, const int* offsets,
                          const int* neighbors, int64_t width, int64_t begin,
                          int64_t end, int64_t* result )
{
    for ( int64_t i = begin; i < end; ++i )
    {
        int nn = counts[i];
        for ( int n = 0; n < nn; ++n )
            result[i] += list_get( layout, offsets, neighbors, width, i, n );
    }
}


// -----------------------------------------------------------------------------
// neighbor_parallel_for directly on a LinkedCellList (no stored list):
// LinkedCellParallelFor (core/src/Cabana_Parallel.hpp:1122-1290) with
// NeighborList<LinkedCellList>::getStencilCells / getParticle semantics
// (core/src/Cabana_LinkedCellList.hpp:840-872): for particle i in [pbegin,pend), walk the
// stencil cells of its bin (decoded on the STENCIL grid, :105-119), candidates
// j = sorted ? offset + lcl_begin : permute[offset]; the functor runs for j != i
// (NeighborDiscriminator<SelfNeighborTag>, Cabana_NeighborList.hpp:68-83) and applies its
// own cutoff.
//   mode 0: result_count[i] += 1 for r2 <= cutoff^2   (tstLinkedCellList.hpp:704-780)
//   mode 1: Lennard-Jones force on i (r2 < cutoff^2), f[3*i+c] accumulated
// -----------------------------------------------------------------------------
void orc_lcl_neighbor_for( const orc_positions* x, const orc_grid* grid,
                           const orc_stencil* stencil, const int* counts,
                           const int64_t* offsets, const int64_t* permute,
                           const int* particle_bins, int sorted, int64_t lcl_begin,
                           int64_t pbegin, int64_t pend, int mode, double cutoff,
                           double eps, double sigma, int* result_count, double* f,
                           double* fabs_out )
{
    const double c2 = cutoff * cutoff;
    const double s2 = sigma * sigma;
#pragma omp parallel for schedule( dynamic, 64 )
    for ( int64_t i = pbegin; i < pend; ++i )
    {
        int mn[3], mx[3];
        orc_stencil_cells( stencil, particle_bins[i - lcl_begin], mn, mx );
        const double xi[3] = { pos_at( x, i, 0 ), pos_at( x, i, 1 ), pos_at( x, i, 2 ) };
        for ( int gi = mn[0]; gi < mx[0]; ++gi )
            for ( int gj = mn[1]; gj < mx[1]; ++gj )
                for ( int gk = mn[2]; gk < mx[2]; ++gk )
                {
                    const int c = cardinal( grid, gi, gj, gk );
                    for ( int64_t n = offsets[c]; n < offsets[c] + counts[c]; ++n )
                    {
                        const int64_t j = sorted ? n + lcl_begin : permute[n];
                        if ( i == j )
                            continue;
                        const double d[3] = { xi[0] - pos_at( x, j, 0 ),
                                              xi[1] - pos_at( x, j, 1 ),
                                              xi[2] - pos_at( x, j, 2 ) };
                        const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                        if ( mode == 0 )
                        {
                            if ( r2 <= c2 )
                                result_count[i] += 1;
                        }
                        else if ( r2 < c2 )
                        {
                            const double sr2 = s2 / r2;
                            const double sr6 = sr2 * sr2 * sr2;
                            const double fpair = 24.0 * eps * sr6 * ( 2.0 * sr6 - 1.0 ) / r2;
                            for ( int k = 0; k < 3; ++k )
                            {
                                f[3 * i + k] += fpair * d[k];
                                if ( fabs_out )
                                    fabs_out[3 * i + k] += std::fabs( fpair * d[k] );
                            }
                        }
                    }
                }
    }
}

int orc_num_threads( void )
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads( int n )
{
#ifdef _OPENMP
    omp_set_num_threads( n );
#else
    (void)n;
#endif
}

} // extern "C"
