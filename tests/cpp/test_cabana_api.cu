// C++ test of include/Cabana_B200.hpp, written the way the reference's own unit tests are
// (core/unit_test/tstNeighborList.hpp, neighbor_unit_test.hpp, tstLinkedCellList.hpp):
// build lists through the Cabana-named classes, compare with a brute-force N^2 list
// computed in the test, exercise neighbor_parallel_for / reduce with user functors.
//
// Compiled with nvcc and linked against libcabana_b200.so by tests/test_gpu_cpp_shim.py.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "Cabana_B200.hpp"

static int g_fail = 0;
#define EXPECT_TRUE( c )                                                       \
    do                                                                         \
    {                                                                          \
        if ( !( c ) )                                                          \
        {                                                                      \
            std::printf( "FAIL %s:%d  %s\n", __FILE__, __LINE__, #c );         \
            ++g_fail;                                                          \
        }                                                                      \
    } while ( 0 )
#define EXPECT_EQ( a, b ) EXPECT_TRUE( ( a ) == ( b ) )

// ---- test data: NeighborListTestData<3> (neighbor_unit_test.hpp:995-1045) ------------------
struct TestData
{
    std::size_t num_particle = 300;
    std::size_t begin = 75, end = 225;
    double test_radius = 2.32;
    double box_min = -5.3 * 2.32, box_max = 4.7 * 2.32;
    double cell_size_ratio = 0.5;
    std::array<double, 3> grid_min, grid_max;
    std::vector<double> xyz; // (n,3)
    int vlen = 32;
    std::size_t stride = 3 * 32;
    double* d_aosoa = nullptr; // AoSoA<MemberTypes<double[3]>> layout on the device

    TestData()
    {
        grid_min = { box_min, box_min, box_min };
        grid_max = { box_max, box_max, box_max };
        xyz.resize( 3 * num_particle );
        std::uint64_t s = 88172645463325252ull; // xorshift64, any fixed seed
        for ( auto& v : xyz )
        {
            s ^= s << 13;
            s ^= s >> 7;
            s ^= s << 17;
            v = box_min + ( box_max - box_min ) * ( ( s >> 11 ) * ( 1.0 / 9007199254740992.0 ) );
        }
        const std::size_t nsoa = ( num_particle + vlen - 1 ) / vlen;
        std::vector<double> host( nsoa * stride, 0.0 );
        for ( std::size_t i = 0; i < num_particle; ++i )
            for ( int d = 0; d < 3; ++d )
                host[stride * ( i / vlen ) + ( i % vlen ) + vlen * d] = xyz[3 * i + d];
        cudaMalloc( &d_aosoa, host.size() * sizeof( double ) );
        cudaMemcpy( d_aosoa, host.data(), host.size() * sizeof( double ),
                    cudaMemcpyHostToDevice );
    }
    ~TestData() { cudaFree( d_aosoa ); }
    Cabana::Slice<double, 3> positions() const
    {
        return Cabana::Slice<double, 3>( d_aosoa, num_particle, stride, vlen );
    }
};

// computeFullNeighborList (neighbor_unit_test.hpp:86-158)
static std::vector<std::vector<int>> bruteForce( const TestData& t )
{
    std::vector<std::vector<int>> nl( t.num_particle );
    const double rsqr = t.test_radius * t.test_radius;
    for ( std::size_t i = 0; i < t.num_particle; ++i )
        for ( std::size_t j = 0; j < t.num_particle; ++j )
            if ( i != j )
            {
                double dsqr = 0.0;
                for ( int d = 0; d < 3; ++d )
                    dsqr += ( t.xyz[3 * i + d] - t.xyz[3 * j + d] ) *
                            ( t.xyz[3 * i + d] - t.xyz[3 * j + d] );
                if ( dsqr <= rsqr )
                    nl[i].push_back( (int)j );
            }
    return nl;
}

template <class ListType>
static std::vector<std::vector<int>> copyListToHost( const ListType& list, std::size_t n )
{
    const cb_verlet_view& v = list.view();
    std::vector<int> counts( n );
    cudaMemcpy( counts.data(), v.counts, n * sizeof( int ), cudaMemcpyDeviceToHost );
    std::vector<std::vector<int>> out( n );
    if ( v.layout == CB_LAYOUT_CSR )
    {
        std::vector<int> offsets( n ), nb( std::max<std::size_t>( v.total, 1 ) );
        cudaMemcpy( offsets.data(), v.offsets, n * sizeof( int ), cudaMemcpyDeviceToHost );
        cudaMemcpy( nb.data(), v.neighbors, v.total * sizeof( int ), cudaMemcpyDeviceToHost );
        for ( std::size_t i = 0; i < n; ++i )
            out[i].assign( nb.begin() + offsets[i], nb.begin() + offsets[i] + counts[i] );
    }
    else
    {
        std::vector<int> nb( std::max<std::size_t>( n * v.width, 1 ) );
        cudaMemcpy( nb.data(), v.neighbors, n * v.width * sizeof( int ), cudaMemcpyDeviceToHost );
        for ( std::size_t i = 0; i < n; ++i )
            out[i].assign( nb.begin() + i * v.width, nb.begin() + i * v.width + counts[i] );
    }
    for ( auto& r : out )
        std::sort( r.begin(), r.end() );
    return out;
}

// checkFullNeighborList (neighbor_unit_test.hpp:161-198)
template <class ListType>
static void checkFullNeighborList( const ListType& list, const TestData& t,
                                   const std::vector<std::vector<int>>& n2, std::size_t b,
                                   std::size_t e )
{
    auto got = copyListToHost( list, t.num_particle );
    std::size_t total = 0, mx = 0;
    for ( std::size_t i = 0; i < t.num_particle; ++i )
    {
        if ( i >= b && i < e )
        {
            EXPECT_TRUE( got[i] == n2[i] );
            total += n2[i].size();
            mx = std::max( mx, n2[i].size() );
        }
        else
            EXPECT_EQ( got[i].size(), 0u );
    }
    EXPECT_EQ( Cabana::NeighborList<ListType>::totalNeighbor( list ), total );
    EXPECT_EQ( Cabana::NeighborList<ListType>::maxNeighbor( list ), mx );
}

// checkHalfNeighborList (neighbor_unit_test.hpp:201-243)
template <class ListType>
static void checkHalfNeighborList( const ListType& list, const TestData& t,
                                   const std::vector<std::vector<int>>& n2 )
{
    auto got = copyListToHost( list, t.num_particle );
    std::size_t full = 0, half = 0;
    for ( std::size_t i = 0; i < t.num_particle; ++i )
    {
        full += n2[i].size();
        half += got[i].size();
        for ( int j : got[i] )
        {
            EXPECT_TRUE( std::binary_search( n2[i].begin(), n2[i].end(), j ) );
            EXPECT_TRUE( !std::binary_search( got[j].begin(), got[j].end(), (int)i ) );
        }
    }
    EXPECT_EQ( full, 2 * half );
    EXPECT_EQ( Cabana::NeighborList<ListType>::totalNeighbor( list ), full / 2 );
}

// ---- user functors (neighbor_unit_test.hpp:291-419) ----------------------------------------
struct IdSumFunctor
{
    long long* result;
    __device__ void operator()( const int i, const int j ) const
    {
        atomicAdd( reinterpret_cast<unsigned long long*>( result + i ), (unsigned long long)j );
    }
};
struct TaggedIdSumFunctor
{
    struct Tag
    {
    };
    long long* result;
    __device__ void operator()( const Tag&, const int i, const int j ) const
    {
        atomicAdd( reinterpret_cast<unsigned long long*>( result + i ),
                   (unsigned long long)( 2 * j ) );
    }
};
// second-neighbour functors (neighbor_unit_test.hpp:474-492, :632-642)
struct TripletIdSumFunctor
{
    long long* result;
    __device__ void operator()( const int i, const int j, const int k ) const
    {
        atomicAdd( reinterpret_cast<unsigned long long*>( result + i ),
                   (unsigned long long)( j + k ) );
    }
};
struct TripletSumPositionsFunctor
{
    Cabana::Slice<double, 3> x;
    __device__ void operator()( const int i, const int j, const int k, double& sum ) const
    {
        sum += x( i, 0 ) + x( j, 0 ) + x( k, 0 );
    }
};
struct IdReduceFunctor
{
    __device__ void operator()( const int, const int j, long long& sum ) const { sum += j; }
};
struct SumPositionsFunctor
{
    Cabana::Slice<double, 3> x;
    __device__ void operator()( const int i, const int j, double& sum ) const
    {
        sum += x( i, 0 ) + x( j, 0 );
    }
};

template <class LayoutTag>
static void testVerletListFull()
{
    TestData t;
    auto n2 = bruteForce( t );
    auto x = t.positions();
    {
        Cabana::VerletList<Cabana::DeviceSpace, Cabana::FullNeighborTag, LayoutTag,
                           Cabana::TeamOpTag>
            nlist( x, 0, x.size(), t.test_radius, t.cell_size_ratio, t.grid_min, t.grid_max );
        checkFullNeighborList( nlist, t, n2, 0, t.num_particle );
        // default construct then assign (tstNeighborList.hpp:42-46)
        Cabana::VerletList<Cabana::DeviceSpace, Cabana::FullNeighborTag, LayoutTag,
                           Cabana::TeamOpTag>
            nlist2;
        nlist2 = nlist;
        checkFullNeighborList( nlist2, t, n2, 0, t.num_particle );
    }
    for ( std::size_t max_neigh : { std::size_t( 100 ), std::size_t( 2 ) } )
    {
        // max_neigh large (no recount) and too small (refill), tstNeighborList.hpp:58-77
        auto nlist = Cabana::createVerletList<Cabana::FullNeighborTag, LayoutTag,
                                              Cabana::TeamVectorOpTag>(
            x, 0, x.size(), t.test_radius, t.cell_size_ratio, t.grid_min, t.grid_max,
            max_neigh );
        checkFullNeighborList( nlist, t, n2, 0, t.num_particle );
    }
    {
        // partial range (tstNeighborList.hpp:112-141)
        Cabana::VerletList<Cabana::DeviceSpace, Cabana::FullNeighborTag, LayoutTag> nlist(
            x, t.begin, t.end, t.test_radius, t.cell_size_ratio, t.grid_min, t.grid_max );
        checkFullNeighborList( nlist, t, n2, t.begin, t.end );
    }
    {
        // rank-2 view instead of a slice (testNeighborView, tstNeighborList.hpp:295-325)
        double* d = nullptr;
        cudaMalloc( &d, t.xyz.size() * sizeof( double ) );
        cudaMemcpy( d, t.xyz.data(), t.xyz.size() * sizeof( double ), cudaMemcpyHostToDevice );
        Cabana::View2D<double, 3> xv( d, t.num_particle );
        double c_min[3] = { t.box_min, t.box_min, t.box_min };
        double c_max[3] = { t.box_max, t.box_max, t.box_max };
        Cabana::VerletList<Cabana::DeviceSpace, Cabana::FullNeighborTag, LayoutTag> nlist(
            xv, 0, xv.size(), t.test_radius, t.cell_size_ratio, c_min, c_max );
        checkFullNeighborList( nlist, t, n2, 0, t.num_particle );
        cudaFree( d );
    }
}

template <class LayoutTag>
static void testVerletListHalf()
{
    TestData t;
    auto n2 = bruteForce( t );
    auto x = t.positions();
    Cabana::VerletList<Cabana::DeviceSpace, Cabana::HalfNeighborTag, LayoutTag> nlist(
        x, 0, x.size(), t.test_radius, t.cell_size_ratio, t.grid_min, t.grid_max );
    checkHalfNeighborList( nlist, t, n2 );
}

template <class LayoutTag>
static void testNeighborParallelFor()
{
    TestData t;
    auto n2 = bruteForce( t );
    auto x = t.positions();
    using ListType = Cabana::VerletList<Cabana::DeviceSpace, Cabana::FullNeighborTag, LayoutTag>;
    ListType nlist( x, 0, x.size(), t.test_radius, t.cell_size_ratio, t.grid_min, t.grid_max );
    const std::size_t n = t.num_particle;
    std::vector<long long> expect( n, 0 );
    double expect_sum = 0.0;
    for ( std::size_t i = 0; i < n; ++i )
        for ( int j : n2[i] )
        {
            expect[i] += j;
            expect_sum += t.xyz[3 * i] + t.xyz[3 * j];
        }
    long long* d_res = nullptr;
    cudaMalloc( &d_res, n * sizeof( long long ) );
    std::vector<long long> res( n );
    Cabana::RangePolicy<> policy( 0, n );

    // Serial and Team, untagged functor
    for ( int team = 0; team < 2; ++team )
    {
        cudaMemset( d_res, 0, n * sizeof( long long ) );
        IdSumFunctor f{ d_res };
        if ( team )
            Cabana::neighbor_parallel_for( policy, f, nlist, Cabana::FirstNeighborsTag(),
                                           Cabana::TeamOpTag(), "test_team" );
        else
            Cabana::neighbor_parallel_for( policy, f, nlist, Cabana::FirstNeighborsTag(),
                                           Cabana::SerialOpTag(), "test_serial" );
        cudaMemcpy( res.data(), d_res, n * sizeof( long long ), cudaMemcpyDeviceToHost );
        EXPECT_TRUE( res == expect );
    }
    // functor with a work tag: multiplier 2 (neighbor_unit_test.hpp:666-742)
    {
        cudaMemset( d_res, 0, n * sizeof( long long ) );
        TaggedIdSumFunctor f{ d_res };
        Cabana::RangePolicy<TaggedIdSumFunctor::Tag> tagged( 0, n );
        Cabana::neighbor_parallel_for( tagged, f, nlist, Cabana::FirstNeighborsTag(),
                                       Cabana::SerialOpTag() );
        cudaMemcpy( res.data(), d_res, n * sizeof( long long ), cudaMemcpyDeviceToHost );
        bool ok = true;
        for ( std::size_t i = 0; i < n; ++i )
            ok = ok && res[i] == 2 * expect[i];
        EXPECT_TRUE( ok );
    }
    // neighbor_parallel_reduce (checkFirstNeighborParallelReduce, EXPECT_FLOAT_EQ)
    for ( int team = 0; team < 2; ++team )
    {
        double sum = 0.0;
        SumPositionsFunctor f{ x };
        if ( team )
            Cabana::neighbor_parallel_reduce( policy, f, nlist, Cabana::FirstNeighborsTag(),
                                              Cabana::TeamOpTag(), sum );
        else
            Cabana::neighbor_parallel_reduce( policy, f, nlist, Cabana::FirstNeighborsTag(),
                                              Cabana::SerialOpTag(), sum );
        EXPECT_TRUE( std::fabs( sum - expect_sum ) <= 1e-6 * std::fabs( expect_sum ) );
    }
    // pre-compiled Lennard-Jones consumer: Serial == Team within 1e-12 of sum |pair force|
    {
        double *d_f = nullptr, *d_g = nullptr;
        cudaMalloc( &d_f, 3 * n * sizeof( double ) );
        cudaMalloc( &d_g, 3 * n * sizeof( double ) );
        cudaMemset( d_f, 0, 3 * n * sizeof( double ) );
        cudaMemset( d_g, 0, 3 * n * sizeof( double ) );
        Cabana::View2D<double, 3> f( d_f, n ), g( d_g, n );
        Cabana::B200::neighbor_parallel_for_lj( policy, nlist, x, f, 1.0, 1.0, 2.0,
                                                Cabana::SerialOpTag() );
        Cabana::B200::neighbor_parallel_for_lj( policy, nlist, x, g, 1.0, 1.0, 2.0,
                                                Cabana::TeamOpTag() );
        std::vector<double> hf( 3 * n ), hg( 3 * n );
        cudaMemcpy( hf.data(), d_f, 3 * n * sizeof( double ), cudaMemcpyDeviceToHost );
        cudaMemcpy( hg.data(), d_g, 3 * n * sizeof( double ), cudaMemcpyDeviceToHost );
        // host reference with the same functor definition
        bool ok = true;
        for ( std::size_t i = 0; i < n; ++i )
        {
            double fr[3] = { 0, 0, 0 }, fa[3] = { 0, 0, 0 };
            for ( int j : n2[i] )
            {
                double d[3], r2 = 0;
                for ( int c = 0; c < 3; ++c )
                {
                    d[c] = t.xyz[3 * i + c] - t.xyz[3 * j + c];
                    r2 += d[c] * d[c];
                }
                if ( r2 < 4.0 )
                {
                    const double sr2 = 1.0 / r2, sr6 = sr2 * sr2 * sr2;
                    const double fp = 24.0 * sr6 * ( 2.0 * sr6 - 1.0 ) / r2;
                    for ( int c = 0; c < 3; ++c )
                    {
                        fr[c] += fp * d[c];
                        fa[c] += std::fabs( fp * d[c] );
                    }
                }
            }
            for ( int c = 0; c < 3; ++c )
            {
                ok = ok && std::fabs( hf[3 * i + c] - fr[c] ) <= 1e-12 * std::max( fa[c], 1e-300 );
                ok = ok && std::fabs( hg[3 * i + c] - fr[c] ) <= 1e-12 * std::max( fa[c], 1e-300 );
            }
        }
        EXPECT_TRUE( ok );
        const double e_s = Cabana::B200::neighbor_parallel_reduce_lj(
            policy, nlist, x, 1.0, 1.0, 2.0, Cabana::SerialOpTag() );
        const double e_t = Cabana::B200::neighbor_parallel_reduce_lj(
            policy, nlist, x, 1.0, 1.0, 2.0, Cabana::TeamOpTag() );
        EXPECT_TRUE( std::fabs( e_s - e_t ) <= 1e-11 * std::fabs( e_s ) );
        cudaFree( d_f );
        cudaFree( d_g );
    }
    // SecondNeighborsTag: Serial, Team, TeamVector (checkSecondNeighborParallelFor,
    // neighbor_unit_test.hpp:327-362) and the reduce (checkSecondNeighborParallelReduce)
    {
        std::vector<long long> expect2( n, 0 );
        double expect_sum2 = 0.0;
        for ( std::size_t i = 0; i < n; ++i )
            for ( std::size_t a = 0; a < n2[i].size(); ++a )
                for ( std::size_t b = a + 1; b < n2[i].size(); ++b )
                {
                    expect2[i] += n2[i][a] + n2[i][b];
                    expect_sum2 += t.xyz[3 * i] + t.xyz[3 * n2[i][a]] + t.xyz[3 * n2[i][b]];
                }
        for ( int op = 0; op < 3; ++op )
        {
            cudaMemset( d_res, 0, n * sizeof( long long ) );
            TripletIdSumFunctor f{ d_res };
            if ( op == 0 )
                Cabana::neighbor_parallel_for( policy, f, nlist, Cabana::SecondNeighborsTag(),
                                               Cabana::SerialOpTag(), "test_2nd_serial" );
            else if ( op == 1 )
                Cabana::neighbor_parallel_for( policy, f, nlist, Cabana::SecondNeighborsTag(),
                                               Cabana::TeamOpTag(), "test_2nd_team" );
            else
                Cabana::neighbor_parallel_for( policy, f, nlist, Cabana::SecondNeighborsTag(),
                                               Cabana::TeamVectorOpTag(), "test_2nd_vector" );
            cudaMemcpy( res.data(), d_res, n * sizeof( long long ), cudaMemcpyDeviceToHost );
            EXPECT_TRUE( res == expect2 );
        }
        for ( int op = 0; op < 2; ++op )
        {
            double sum = 0.0;
            TripletSumPositionsFunctor f{ x };
            if ( op == 0 )
                Cabana::neighbor_parallel_reduce( policy, f, nlist,
                                                  Cabana::SecondNeighborsTag(),
                                                  Cabana::SerialOpTag(), sum );
            else
                Cabana::neighbor_parallel_reduce( policy, f, nlist,
                                                  Cabana::SecondNeighborsTag(),
                                                  Cabana::TeamOpTag(), sum );
            EXPECT_TRUE( std::fabs( sum - expect_sum2 ) <= 1e-6 * std::fabs( expect_sum2 ) );
        }
    }
    // setNeighbor (testModifyNeighbors, tstNeighborList.hpp:256-292)
    {
        nlist.setNeighbor( 3, 0, -5 );
        EXPECT_EQ( Cabana::NeighborList<ListType>::getNeighbor( nlist, 3, 0 ), (std::size_t)-5 );
    }
    cudaFree( d_res );
}

// testLinkedList (tstLinkedCellList.hpp:584-620): one particle per unit cell, created
// x-fastest; after bin + permute the order is i-slowest / k-fastest.
static void testLinkedList()
{
    const int nx = 10;
    const std::size_t n = nx * nx * nx;
    std::vector<double> xyz( 3 * n );
    std::vector<int> ids( 3 * n );
    for ( int k = 0; k < nx; ++k )
        for ( int j = 0; j < nx; ++j )
            for ( int i = 0; i < nx; ++i )
            {
                const std::size_t p = i + j * nx + k * nx * nx;
                xyz[3 * p + 0] = i + 0.5;
                xyz[3 * p + 1] = j + 0.5;
                xyz[3 * p + 2] = k + 0.5;
                ids[3 * p + 0] = i;
                ids[3 * p + 1] = j;
                ids[3 * p + 2] = k;
            }
    double* d_x = nullptr;
    int* d_id = nullptr;
    cudaMalloc( &d_x, xyz.size() * sizeof( double ) );
    cudaMalloc( &d_id, ids.size() * sizeof( int ) );
    cudaMemcpy( d_x, xyz.data(), xyz.size() * sizeof( double ), cudaMemcpyHostToDevice );
    cudaMemcpy( d_id, ids.data(), ids.size() * sizeof( int ), cudaMemcpyHostToDevice );
    Cabana::View2D<double, 3> pos( d_x, n );
    Cabana::View2D<int, 3> cell_id( d_id, n );
    std::array<double, 3> delta = { 1, 1, 1 }, mn = { 0, 0, 0 }, mx = { 10, 10, 10 };
    for ( int partial = 0; partial < 2; ++partial )
    {
        cudaMemcpy( d_x, xyz.data(), xyz.size() * sizeof( double ), cudaMemcpyHostToDevice );
        cudaMemcpy( d_id, ids.data(), ids.size() * sizeof( int ), cudaMemcpyHostToDevice );
        const std::size_t b = partial ? 250 : 0, e = partial ? 750 : n;
        auto lcl = partial ? Cabana::createLinkedCellList( pos, b, e, delta, mn, mx )
                           : Cabana::createLinkedCellList( pos, delta, mn, mx );
        EXPECT_EQ( lcl.totalBins(), 1000 );
        EXPECT_EQ( lcl.numBin( 0 ), 10 );
        EXPECT_TRUE( !lcl.sorted() );
        Cabana::permute( lcl, pos, cell_id );
        EXPECT_TRUE( lcl.sorted() );
        auto m = lcl.hostMirror();
        std::vector<int> hid( 3 * n );
        cudaMemcpy( hid.data(), d_id, hid.size() * sizeof( int ), cudaMemcpyDeviceToHost );
        // checkLinkedCell (tstLinkedCellList.hpp:281-365)
        std::size_t particle_id = 0;
        bool ok = true;
        for ( int i = 0; i < nx; ++i )
            for ( int j = 0; j < nx; ++j )
                for ( int k = 0; k < nx; ++k )
                {
                    const std::size_t original = i + j * nx + k * nx * nx;
                    const auto c = lcl.cardinalBinIndex( i, j, k );
                    if ( b <= original && original < e )
                    {
                        const std::size_t sort_id = b + particle_id;
                        ok = ok && hid[3 * sort_id] == i && hid[3 * sort_id + 1] == j &&
                             hid[3 * sort_id + 2] == k;
                        ok = ok && m.counts[c] == 1 && m.offsets[c] == particle_id;
                        ++particle_id;
                    }
                    else
                        ok = ok && m.counts[c] == 0;
                }
        for ( std::size_t p = 0; p < n; ++p )
            if ( p < b || p >= e )
                ok = ok && hid[3 * p] == (int)( p % nx ) && hid[3 * p + 1] == (int)( ( p / nx ) % nx ) &&
                     hid[3 * p + 2] == (int)( p / ( nx * nx ) );
        EXPECT_TRUE( ok );
        int imin, imax, jmin, jmax, kmin, kmax;
        lcl.getStencilCells( (int)lcl.cardinalBinIndex( 4, 5, 3 ), imin, imax, jmin, jmax, kmin,
                             kmax );
        EXPECT_TRUE( imin == 3 && imax == 6 && jmin == 4 && jmax == 7 && kmin == 2 && kmax == 5 );
    }
    cudaFree( d_x );
    cudaFree( d_id );
}

// Device-side LinkedCellList accessors used inside a user kernel (the reference's functors call
// binSize/binOffset/permutation/getParticleBin/getStencilCells on the captured list), and
// permute(BinningData, slices) (tstSort.hpp / Cabana_Sort.hpp:549-715).
__global__ void k_check_lcl_view( Cabana::LinkedCellListView l, int n, int* bad )
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if ( p >= n )
        return;
    // slot p of the binned order holds particle permutation(p); its bin must contain slot p
    const int pid = (int)l.permutation( p );
    // after permute() the particle in slot p IS particle rangeBegin()+p (storeParticleBins, :806-823)
    const int cell = l.getParticleBin( l.sorted() ? p + (int)l.rangeBegin() : pid );
    int i, j, k;
    l.ijkBinIndex( cell, i, j, k );
    const unsigned off = l.binOffset( i, j, k );
    const int size = l.binSize( i, j, k );
    if ( l.cardinalBinIndex( i, j, k ) != cell || (unsigned)p < off || (unsigned)p >= off + size )
        atomicAdd( bad, 1 );
    if ( (int)l.getParticle( p ) != ( l.sorted() ? p + (int)l.rangeBegin() : pid ) )
        atomicAdd( bad, 1 );
    int imin, imax, jmin, jmax, kmin, kmax;
    l.getStencilCells( cell, imin, imax, jmin, jmax, kmin, kmax );
    if ( i < imin || i >= imax || j < jmin || j >= jmax || k < kmin || k >= kmax )
        atomicAdd( bad, 1 );
}

// NeighborList<LinkedCellList> traits on the device (Cabana_LinkedCellList.hpp:1149-1303)
__global__ void k_lcl_traits( Cabana::LinkedCellListView l, int n, long long* nn, long long* idsum )
{
    using traits = Cabana::NeighborList<Cabana::LinkedCellListView>;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if ( p >= n )
        return;
    const std::size_t c = traits::numNeighbor( l, p );
    long long sum = 0;
    for ( std::size_t k = 0; k < c; ++k )
        sum += (long long)traits::getNeighbor( l, p, k );
    nn[p] = (long long)c;
    idsum[p] = sum;
}

static void testBinningData()
{
    TestData t;
    const std::size_t n = t.num_particle;
    double *d_x = nullptr, *d_y = nullptr;
    int* d_bad = nullptr;
    cudaMalloc( &d_x, 3 * n * sizeof( double ) );
    cudaMalloc( &d_y, 3 * n * sizeof( double ) );
    cudaMalloc( &d_bad, sizeof( int ) );
    cudaMemset( d_bad, 0, sizeof( int ) );
    cudaMemcpy( d_x, t.xyz.data(), 3 * n * sizeof( double ), cudaMemcpyHostToDevice );
    cudaMemcpy( d_y, t.xyz.data(), 3 * n * sizeof( double ), cudaMemcpyHostToDevice );
    Cabana::View2D<double, 3> pos( d_x, n ), pos2( d_y, n );
    const double dcell = t.test_radius * t.cell_size_ratio;
    std::array<double, 3> delta = { dcell, dcell, dcell };
    auto lcl = Cabana::createLinkedCellList( pos, delta, t.grid_min, t.grid_max, t.test_radius,
                                             t.cell_size_ratio );
    auto bin_data = lcl.binningData();
    EXPECT_EQ( bin_data.numBin(), lcl.totalBins() );
    EXPECT_EQ( bin_data.rangeBegin(), std::size_t( 0 ) );
    EXPECT_EQ( bin_data.rangeEnd(), n );
    k_check_lcl_view<<<( (int)n + 127 ) / 128, 128>>>( lcl.deviceView(), (int)n, d_bad );
    // permute through the binning data == permute through the list
    Cabana::permute( bin_data, pos2 );
    Cabana::permute( lcl, pos );
    k_check_lcl_view<<<( (int)n + 127 ) / 128, 128>>>( lcl.deviceView(), (int)n, d_bad );
    std::vector<double> a( 3 * n ), b( 3 * n );
    cudaMemcpy( a.data(), d_x, 3 * n * sizeof( double ), cudaMemcpyDeviceToHost );
    cudaMemcpy( b.data(), d_y, 3 * n * sizeof( double ), cudaMemcpyDeviceToHost );
    EXPECT_TRUE( a == b );
    auto m = lcl.hostMirror();
    bool ok = true;
    for ( std::size_t p = 0; p < n; ++p )
        for ( int d = 0; d < 3; ++d )
            ok = ok && a[3 * p + d] == t.xyz[3 * m.permute[p] + d];
    EXPECT_TRUE( ok );
    int bad = -1;
    cudaMemcpy( &bad, d_bad, sizeof( int ), cudaMemcpyDeviceToHost );
    EXPECT_EQ( bad, 0 );
    // traits: every particle "sees" all particles of its stencil cells (itself included)
    {
        long long *d_nn = nullptr, *d_sum = nullptr;
        cudaMalloc( &d_nn, n * sizeof( long long ) );
        cudaMalloc( &d_sum, n * sizeof( long long ) );
        k_lcl_traits<<<( (int)n + 127 ) / 128, 128>>>( lcl.deviceView(), (int)n, d_nn, d_sum );
        std::vector<long long> nn( n ), sum( n );
        cudaMemcpy( nn.data(), d_nn, n * sizeof( long long ), cudaMemcpyDeviceToHost );
        cudaMemcpy( sum.data(), d_sum, n * sizeof( long long ), cudaMemcpyDeviceToHost );
        // host replay from the mirror (the list is sorted now: slot p holds particle p)
        bool ok2 = true;
        long long total = 0;
        for ( std::size_t p = 0; p < n; ++p )
        {
            int imin, imax, jmin, jmax, kmin, kmax;
            lcl.getStencilCells( m.particle_bins[p], imin, imax, jmin, jmax, kmin, kmax );
            long long c = 0, sid = 0;
            for ( int i = imin; i < imax; ++i )
                for ( int j = jmin; j < jmax; ++j )
                    for ( int k = kmin; k < kmax; ++k )
                    {
                        const auto cell = lcl.cardinalBinIndex( i, j, k );
                        for ( int q = 0; q < m.counts[cell]; ++q )
                            sid += (long long)( m.offsets[cell] + q ); // sorted: id = slot
                        c += m.counts[cell];
                    }
            ok2 = ok2 && nn[p] == c && sum[p] == sid;
            total += c;
        }
        EXPECT_TRUE( ok2 );
        EXPECT_TRUE( total > (long long)n );
        cudaFree( d_nn );
        cudaFree( d_sum );
    }
    cudaFree( d_x );
    cudaFree( d_y );
    cudaFree( d_bad );
}

// testNeighborHistogram (tstNeighborList.hpp:328-381): 10^3 lattice, r = 3 dx + 1e-7; the
// histogram values are the reference's literal known answers.
static void testNeighborHistogram()
{
    const int np = 10;
    const std::size_t n = np * np * np;
    const double dx = 5.0 / np;
    std::vector<double> xyz( 3 * n );
    for ( std::size_t pid = 0; pid < n; ++pid )
    {
        xyz[3 * pid + 0] = dx / 2 + dx * ( pid / ( np * np ) );
        xyz[3 * pid + 1] = dx / 2 + dx * ( ( pid / np ) % np );
        xyz[3 * pid + 2] = dx / 2 + dx * ( pid % np );
    }
    double* d_x = nullptr;
    cudaMalloc( &d_x, xyz.size() * sizeof( double ) );
    cudaMemcpy( d_x, xyz.data(), xyz.size() * sizeof( double ), cudaMemcpyHostToDevice );
    Cabana::View2D<double, 3> pos( d_x, n );
    std::array<double, 3> mn = { 0, 0, 0 }, mx = { 5, 5, 5 };
    using ListType =
        Cabana::VerletList<Cabana::DeviceSpace, Cabana::FullNeighborTag, Cabana::VerletLayoutCSR>;
    ListType nlist( pos, 0, n, 3 * dx + 1e-7, 0.5, mn, mx );
    EXPECT_EQ( Cabana::NeighborList<ListType>::maxNeighbor( nlist ), std::size_t( 122 ) );
    {
        auto h = Cabana::neighborHistogram( n, nlist, 10 );
        const int bin_max[10] = { 12, 24, 36, 48, 61, 73, 85, 97, 109, 122 };
        const int bin_count[10] = { 32, 72, 24, 152, 120, 168, 0, 216, 0, 152 };
        for ( int i = 0; i < 10; ++i )
        {
            EXPECT_EQ( h[i][0], bin_max[i] );
            EXPECT_EQ( h[i][1], bin_count[i] );
        }
    }
    {
        auto h = Cabana::neighborHistogram( n, nlist, 5 );
        const int bin_max[5] = { 24, 48, 73, 97, 122 };
        const int bin_count[5] = { 104, 176, 288, 216, 152 };
        for ( int i = 0; i < 5; ++i )
        {
            EXPECT_EQ( h[i][0], bin_max[i] );
            EXPECT_EQ( h[i][1], bin_count[i] );
        }
    }
    cudaFree( d_x );
}

// neighbor_parallel_for directly on a LinkedCellList (tstLinkedCellList.hpp:704-780,
// checkLinkedCellNeighborPar): the functor applies the cutoff; counts must equal the N^2
// list's, Serial and Team, before and after permute.
struct LclCountFunctor
{
    Cabana::View2D<double, 3> x;
    double rsqr;
    int* result;
    __device__ void operator()( const int i, const int j ) const
    {
        const double dx = x( i, 0 ) - x( j, 0 ), dy = x( i, 1 ) - x( j, 1 ),
                     dz = x( i, 2 ) - x( j, 2 );
        const double d2 = __dadd_rn( __dadd_rn( __dmul_rn( dx, dx ), __dmul_rn( dy, dy ) ),
                                     __dmul_rn( dz, dz ) );
        if ( d2 <= rsqr )
            atomicAdd( result + i, 1 );
    }
};

static void testLinkedCellParallelFor()
{
    TestData t;
    auto n2 = bruteForce( t );
    const std::size_t n = t.num_particle;
    double* d_x = nullptr;
    int *d_id = nullptr, *d_res = nullptr;
    cudaMalloc( &d_x, 3 * n * sizeof( double ) );
    cudaMalloc( &d_id, n * sizeof( int ) );
    cudaMalloc( &d_res, n * sizeof( int ) );
    cudaMemcpy( d_x, t.xyz.data(), 3 * n * sizeof( double ), cudaMemcpyHostToDevice );
    std::vector<int> ids( n );
    for ( std::size_t i = 0; i < n; ++i )
        ids[i] = (int)i;
    cudaMemcpy( d_id, ids.data(), n * sizeof( int ), cudaMemcpyHostToDevice );
    Cabana::View2D<double, 3> pos( d_x, n );
    Cabana::View2D<int, 1> pid( d_id, n );
    const double dcell = t.test_radius * t.cell_size_ratio;
    std::array<double, 3> delta = { dcell, dcell, dcell };
    auto lcl = Cabana::createLinkedCellList( pos, delta, t.grid_min, t.grid_max, t.test_radius,
                                             t.cell_size_ratio );
    Cabana::RangePolicy<> policy( 0, n );
    std::vector<int> res( n ), hid( n );
    for ( int sorted = 0; sorted < 2; ++sorted )
    {
        if ( sorted )
        {
            Cabana::permute( lcl, pos, pid );
            EXPECT_TRUE( lcl.sorted() );
        }
        cudaMemcpy( hid.data(), d_id, n * sizeof( int ), cudaMemcpyDeviceToHost );
        for ( int team = 0; team < 2; ++team )
        {
            cudaMemset( d_res, 0, n * sizeof( int ) );
            LclCountFunctor f{ pos, t.test_radius * t.test_radius, d_res };
            if ( team )
                Cabana::neighbor_parallel_for( policy, f, lcl, Cabana::FirstNeighborsTag(),
                                               Cabana::TeamOpTag(), "lcl_team" );
            else
                Cabana::neighbor_parallel_for( policy, f, lcl, Cabana::FirstNeighborsTag(),
                                               Cabana::SerialOpTag(), "lcl_serial" );
            cudaMemcpy( res.data(), d_res, n * sizeof( int ), cudaMemcpyDeviceToHost );
            bool ok = true;
            for ( std::size_t i = 0; i < n; ++i )
                ok = ok && res[i] == (int)n2[hid[i]].size();
            EXPECT_TRUE( ok );
        }
    }
    cudaFree( d_x );
    cudaFree( d_id );
    cudaFree( d_res );
}

// ---- testNonUniformRadius (tstNeighborList.hpp:210-253): per-particle cutoff radius ---------
template <class LayoutTag>
static void testNonUniformRadius()
{
    const int px = 2, n = px * px * px;
    const double dx = 5.0 / px, large_radius = 4.05, small_radius = 3.32;
    std::vector<double> xyz( 3 * n ), rad( n, small_radius );
    for ( int p = 0; p < n; ++p )
    {
        xyz[3 * p + 0] = dx / 2 + dx * ( p / ( px * px ) );
        xyz[3 * p + 1] = dx / 2 + dx * ( ( p / px ) % px );
        xyz[3 * p + 2] = dx / 2 + dx * ( p % px );
    }
    rad[0] = rad[n - 1] = large_radius;
    double *d_x = nullptr, *d_r = nullptr;
    cudaMalloc( &d_x, xyz.size() * sizeof( double ) );
    cudaMalloc( &d_r, rad.size() * sizeof( double ) );
    cudaMemcpy( d_x, xyz.data(), xyz.size() * sizeof( double ), cudaMemcpyHostToDevice );
    cudaMemcpy( d_r, rad.data(), rad.size() * sizeof( double ), cudaMemcpyHostToDevice );
    Cabana::View2D<double, 3> position( d_x, n );
    Cabana::View2D<double, 1> radii( d_r, n );
    std::array<double, 3> grid_min = { 0.0, 0.0, 0.0 }, grid_max = { 5.0, 5.0, 5.0 };
    using ListType = Cabana::VerletList<Cabana::DeviceSpace, Cabana::FullNeighborTag, LayoutTag,
                                        Cabana::TeamOpTag>;
    ListType nlist( position, 0, position.size(), small_radius, radii, 0.5, grid_min, grid_max );
    auto rows = copyListToHost( nlist, n );
    for ( int p = 0; p < n; ++p )
        EXPECT_EQ( (int)rows[p].size(), ( p == 0 || p == n - 1 ) ? 6 : 4 );
    // the same through createVerletList and an explicit execution space (a non-default stream)
    cudaStream_t st;
    cudaStreamCreate( &st );
    Cabana::DeviceExecutionSpace space( st );
    ListType on_stream( space, position, 0, position.size(), small_radius, radii, 0.5, grid_min,
                        grid_max );
    space.fence();
    auto rows2 = copyListToHost( on_stream, n );
    EXPECT_TRUE( rows2 == rows );
    auto made = Cabana::createVerletList<Cabana::FullNeighborTag, LayoutTag, Cabana::TeamOpTag>(
        position, 0, position.size(), small_radius, radii, 0.5, grid_min, grid_max );
    EXPECT_TRUE( copyListToHost( made, n ) == rows );
    cudaStreamDestroy( st );
    cudaFree( d_x );
    cudaFree( d_r );
}

// ---- ExecutionSpace overloads (Cabana_VerletList.hpp:1260-1281, :1351-1392) ----------------
static void testExecutionSpaceOverloads()
{
    TestData t;
    auto n2 = bruteForce( t );
    cudaStream_t st;
    cudaStreamCreateWithFlags( &st, cudaStreamNonBlocking );
    Cabana::DeviceExecutionSpace space( st );
    using ListType = Cabana::VerletList<Cabana::DeviceSpace, Cabana::FullNeighborTag,
                                        Cabana::VerletLayoutCSR, Cabana::TeamVectorOpTag>;
    ListType list;
    list.build( space, t.positions(), 0, t.num_particle, t.test_radius, t.cell_size_ratio,
                t.grid_min, t.grid_max );
    auto made = Cabana::createVerletList<Cabana::FullNeighborTag, Cabana::VerletLayoutCSR,
                                         Cabana::TeamVectorOpTag>(
        space, t.positions(), 0, t.num_particle, t.test_radius, t.cell_size_ratio, t.grid_min,
        t.grid_max );
    space.fence();
    checkFullNeighborList( list, t, n2, 0, t.num_particle );
    checkFullNeighborList( made, t, n2, 0, t.num_particle );
    // traversal on the same stream: RangePolicy( space, begin, end )
    long long* d_sum = nullptr;
    cudaMalloc( &d_sum, t.num_particle * sizeof( long long ) );
    cudaMemsetAsync( d_sum, 0, t.num_particle * sizeof( long long ), st );
    Cabana::RangePolicy<> policy( space, 0, t.num_particle );
    Cabana::neighbor_parallel_for( policy, IdSumFunctor{ d_sum }, list,
                                   Cabana::FirstNeighborsTag(), Cabana::TeamOpTag(), "on_stream" );
    long long total = 0;
    Cabana::neighbor_parallel_reduce( policy, IdReduceFunctor{}, list,
                                      Cabana::FirstNeighborsTag(), Cabana::SerialOpTag(), total );
    space.fence();
    std::vector<long long> sums( t.num_particle );
    cudaMemcpy( sums.data(), d_sum, sums.size() * sizeof( long long ), cudaMemcpyDeviceToHost );
    long long expect_total = 0;
    bool ok = true;
    for ( std::size_t i = 0; i < t.num_particle; ++i )
    {
        long long e = 0;
        for ( int j : n2[i] )
            e += j;
        ok = ok && sums[i] == e;
        expect_total += e;
    }
    EXPECT_TRUE( ok );
    EXPECT_EQ( total, expect_total );
    cudaFree( d_sum );
    cudaStreamDestroy( st );
}

// ---- for_each_neighbor, serial and team forms (Cabana_Parallel.hpp:1058-1110;
//      neighbor_unit_test.hpp:511-590) inside a user kernel ---------------------------------
struct AtomicIdSum
{
    long long* out;
    __device__ void operator()( const int i, const int j ) const
    {
        atomicAdd( reinterpret_cast<unsigned long long*>( out + i ), (unsigned long long)j );
    }
};
template <class ViewType>
__global__ void k_for_each_serial( ViewType list, AtomicIdSum f, int n )
{
    for ( int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x )
        Cabana::for_each_neighbor( i, f, list, Cabana::FirstNeighborsTag() );
}
template <class ViewType>
__global__ void k_for_each_team( ViewType list, AtomicIdSum f, int n )
{
    const Cabana::WarpTeamMember team;
    for ( int i = team.league_rank(); i < n; i += team.league_size() )
        Cabana::for_each_neighbor( i, team, f, list, Cabana::FirstNeighborsTag() );
}
template <class LayoutTag>
static void testForEachNeighbor()
{
    TestData t;
    auto n2 = bruteForce( t );
    using ListType = Cabana::VerletList<Cabana::DeviceSpace, Cabana::FullNeighborTag, LayoutTag,
                                        Cabana::TeamOpTag>;
    ListType list( t.positions(), 0, t.num_particle, t.test_radius, t.cell_size_ratio,
                   t.grid_min, t.grid_max );
    const int n = (int)t.num_particle;
    long long* d_sum = nullptr;
    cudaMalloc( &d_sum, n * sizeof( long long ) );
    std::vector<long long> sums( n );
    for ( int team = 0; team < 2; ++team )
    {
        cudaMemset( d_sum, 0, n * sizeof( long long ) );
        if ( team )
            k_for_each_team<<<16, 128>>>( list.deviceView(), AtomicIdSum{ d_sum }, n );
        else
            k_for_each_serial<<<4, 128>>>( list.deviceView(), AtomicIdSum{ d_sum }, n );
        EXPECT_TRUE( cudaDeviceSynchronize() == cudaSuccess );
        cudaMemcpy( sums.data(), d_sum, n * sizeof( long long ), cudaMemcpyDeviceToHost );
        bool ok = true;
        for ( int i = 0; i < n; ++i )
        {
            long long e = 0;
            for ( int j : n2[i] )
                e += j;
            ok = ok && sums[i] == e;
        }
        EXPECT_TRUE( ok );
    }
    cudaFree( d_sum );
}

// ---- neighbor_parallel_reduce on a LinkedCellList (Cabana_Parallel.hpp:1596-1726;
//      tstLinkedCellList.hpp:704-1003) -------------------------------------------------------
struct LclPairCountReduce
{
    Cabana::View2D<double, 3> x;
    double rsqr;
    __device__ void operator()( const int i, const int j, long long& sum ) const
    {
        const double dx = x( i, 0 ) - x( j, 0 ), dy = x( i, 1 ) - x( j, 1 ),
                     dz = x( i, 2 ) - x( j, 2 );
        const double d2 = __dadd_rn( __dadd_rn( __dmul_rn( dx, dx ), __dmul_rn( dy, dy ) ),
                                     __dmul_rn( dz, dz ) );
        if ( d2 <= rsqr )
            sum += 1 + j; // depends on the neighbour, so a wrong pairing cannot cancel
    }
};
struct LclDistanceSumReduce
{
    Cabana::View2D<double, 3> x;
    double rsqr;
    __device__ void operator()( const int i, const int j, double& sum ) const
    {
        const double dx = x( i, 0 ) - x( j, 0 ), dy = x( i, 1 ) - x( j, 1 ),
                     dz = x( i, 2 ) - x( j, 2 );
        const double d2 = dx * dx + dy * dy + dz * dz;
        if ( d2 <= rsqr )
            sum += d2;
    }
};
static void testLinkedCellParallelReduce()
{
    TestData t;
    auto n2 = bruteForce( t );
    const std::size_t n = t.num_particle;
    double* d_x = nullptr;
    cudaMalloc( &d_x, 3 * n * sizeof( double ) );
    cudaMemcpy( d_x, t.xyz.data(), 3 * n * sizeof( double ), cudaMemcpyHostToDevice );
    Cabana::View2D<double, 3> pos( d_x, n );
    const double dcell = t.test_radius * t.cell_size_ratio;
    std::array<double, 3> delta = { dcell, dcell, dcell };
    auto lcl = Cabana::createLinkedCellList( pos, delta, t.grid_min, t.grid_max, t.test_radius,
                                             t.cell_size_ratio );
    long long expect = 0;
    double expect_d2 = 0.0;
    for ( std::size_t i = 0; i < n; ++i )
        for ( int j : n2[i] )
        {
            expect += 1 + j;
            double d2 = 0.0;
            for ( int d = 0; d < 3; ++d )
                d2 += ( t.xyz[3 * i + d] - t.xyz[3 * j + d] ) * ( t.xyz[3 * i + d] - t.xyz[3 * j + d] );
            expect_d2 += d2;
        }
    Cabana::RangePolicy<> policy( 0, n );
    const double rsqr = t.test_radius * t.test_radius;
    long long got = -1;
    Cabana::neighbor_parallel_reduce( policy, LclPairCountReduce{ pos, rsqr }, lcl,
                                      Cabana::FirstNeighborsTag(), Cabana::SerialOpTag(), got,
                                      "lcl_reduce_serial" );
    EXPECT_EQ( got, expect );
    got = -1;
    Cabana::neighbor_parallel_reduce( policy, LclPairCountReduce{ pos, rsqr }, lcl,
                                      Cabana::FirstNeighborsTag(), Cabana::TeamOpTag(), got,
                                      "lcl_reduce_team" );
    EXPECT_EQ( got, expect );
    double got_d2 = 0.0;
    Cabana::neighbor_parallel_reduce( policy, LclDistanceSumReduce{ pos, rsqr }, lcl,
                                      Cabana::FirstNeighborsTag(), Cabana::TeamOpTag(), got_d2 );
    EXPECT_TRUE( std::fabs( got_d2 - expect_d2 ) <= 1e-10 * expect_d2 );
    // a range that was not binned is refused (the reference asserts, :1651-1652)
    bool threw = false;
    try
    {
        Cabana::RangePolicy<> part( 10, n - 10 );
        Cabana::neighbor_parallel_reduce( part, LclPairCountReduce{ pos, rsqr }, lcl,
                                          Cabana::FirstNeighborsTag(), Cabana::SerialOpTag(), got );
    }
    catch ( const std::runtime_error& )
    {
        threw = true;
    }
    EXPECT_TRUE( threw );
    cudaFree( d_x );
}

// ---- NumSpaceDim = 2 (tstNeighborList.hpp testVerletListFull<2,...>) -------------------------
template <class LayoutTag>
static void testVerletList2d()
{
    const std::size_t n = 400;
    const double r = 2.32, lo = -5.3 * 2.32, hi = 4.7 * 2.32;
    std::vector<double> xy( 2 * n );
    std::uint64_t s = 1234567890123ull;
    for ( auto& v : xy )
    {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        v = lo + ( hi - lo ) * ( ( s >> 11 ) * ( 1.0 / 9007199254740992.0 ) );
    }
    double* d_xy = nullptr;
    cudaMalloc( &d_xy, xy.size() * sizeof( double ) );
    cudaMemcpy( d_xy, xy.data(), xy.size() * sizeof( double ), cudaMemcpyHostToDevice );
    Cabana::View2D<double, 2> pos( d_xy, n );
    std::array<double, 2> gmin = { lo, lo }, gmax = { hi, hi };
    using Full = Cabana::VerletList<Cabana::DeviceSpace, Cabana::FullNeighborTag, LayoutTag,
                                    Cabana::TeamOpTag, 2>;
    using Half = Cabana::VerletList<Cabana::DeviceSpace, Cabana::HalfNeighborTag, LayoutTag,
                                    Cabana::TeamOpTag, 2>;
    Full full( pos, 0, n, r, 0.5, gmin, gmax );
    Half half( pos, 0, n, r, 0.5, gmin, gmax );
    auto rf = copyListToHost( full, n );
    auto rh = copyListToHost( half, n );
    std::size_t tf = 0, th = 0;
    bool ok = true;
    for ( std::size_t i = 0; i < n; ++i )
    {
        std::vector<int> ef, eh;
        for ( std::size_t j = 0; j < n; ++j )
        {
            if ( i == j )
                continue;
            const double dx = xy[2 * i] - xy[2 * j], dy = xy[2 * i + 1] - xy[2 * j + 1];
            if ( dx * dx + dy * dy <= r * r )
            {
                ef.push_back( (int)j );
                if ( xy[2 * j] > xy[2 * i] ||
                     ( xy[2 * j] == xy[2 * i] && xy[2 * j + 1] > xy[2 * i + 1] ) )
                    eh.push_back( (int)j );
            }
        }
        ok = ok && rf[i] == ef && rh[i] == eh;
        tf += ef.size();
        th += eh.size();
    }
    EXPECT_TRUE( ok );
    EXPECT_EQ( tf, 2 * th );
    EXPECT_EQ( Cabana::NeighborList<Full>::totalNeighbor( full ), tf );
    cudaFree( d_xy );
}

int main()
{
    if ( cb_device_count() < 1 )
    {
        std::printf( "no CUDA device\n" );
        return 2;
    }
    testLinkedList();
    testVerletListFull<Cabana::VerletLayoutCSR>();
    testVerletListFull<Cabana::VerletLayout2D>();
    testVerletListHalf<Cabana::VerletLayoutCSR>();
    testVerletListHalf<Cabana::VerletLayout2D>();
    testNeighborParallelFor<Cabana::VerletLayoutCSR>();
    testNeighborParallelFor<Cabana::VerletLayout2D>();
    testLinkedCellParallelFor();
    testLinkedCellParallelReduce();
    testNonUniformRadius<Cabana::VerletLayoutCSR>();
    testNonUniformRadius<Cabana::VerletLayout2D>();
    testExecutionSpaceOverloads();
    testForEachNeighbor<Cabana::VerletLayoutCSR>();
    testForEachNeighbor<Cabana::VerletLayout2D>();
    testVerletList2d<Cabana::VerletLayoutCSR>();
    testVerletList2d<Cabana::VerletLayout2D>();
    testBinningData();
    testNeighborHistogram();
    cudaDeviceSynchronize();
    if ( g_fail == 0 )
        std::printf( "ALL CABANA API TESTS PASSED\n" );
    return g_fail == 0 ? 0 : 1;
}
