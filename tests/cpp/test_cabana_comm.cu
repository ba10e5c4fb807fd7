// C++ test of include/Cabana_B200_Comm.hpp: Halo / Distributor / gather / scatter / migrate over
// NCCL, written after the reference's own tests (core/unit_test/tstHalo.hpp:28-330 -- the
// "unique" and "all" halos, with and without topology -- and tstDistributor.hpp test1/2/4 plus a
// ring), one process per GPU.
//
//   test_cabana_comm.bin <rank> <world> <id-file>
// rank 0 writes the ncclUniqueId to <id-file>; the other ranks wait for it.  world = 1 runs the
// same cases with this rank as its only neighbour (a one-GPU box).
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <thread>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>

#include "Cabana_B200_Comm.hpp"

static int g_fail = 0;
static int g_rank = 0;
#define EXPECT_TRUE( c )                                                                        \
    do                                                                                          \
    {                                                                                           \
        if ( !( c ) )                                                                           \
        {                                                                                       \
            std::printf( "FAIL rank %d %s:%d  %s\n", g_rank, __FILE__, __LINE__, #c );          \
            ++g_fail;                                                                           \
        }                                                                                       \
    } while ( 0 )
#define EXPECT_EQ( a, b ) EXPECT_TRUE( ( a ) == ( b ) )

template <class T>
struct DeviceArray
{
    T* p = nullptr;
    std::size_t n = 0;
    explicit DeviceArray( std::size_t n_ )
        : n( n_ )
    {
        cudaMalloc( &p, sizeof( T ) * ( n ? n : 1 ) );
        cudaMemset( p, 0, sizeof( T ) * ( n ? n : 1 ) );
    }
    DeviceArray( const std::vector<T>& h )
        : DeviceArray( h.size() )
    {
        cudaMemcpy( p, h.data(), sizeof( T ) * n, cudaMemcpyHostToDevice );
    }
    ~DeviceArray() { cudaFree( p ); }
    std::vector<T> host() const
    {
        std::vector<T> h( n );
        cudaMemcpy( h.data(), p, sizeof( T ) * n, cudaMemcpyDeviceToHost );
        return h;
    }
};
using IntView = Cabana::View2D<int, 1>;
using UIntView = Cabana::View2D<unsigned, 1>;
using DblView2 = Cabana::View2D<double, 2>;
using HaloT = Cabana::Halo<Cabana::DeviceSpace, Cabana::Export, Cabana::Nccl>;
using DistT = Cabana::Distributor<Cabana::DeviceSpace, Cabana::Nccl>;

// ---- tstHalo "unique": every rank sends element 2n+1 to rank n (itself included) -----------
static void testHaloUnique( const Cabana::NcclCommunicator& comm, bool use_topology )
{
    const int size = comm.size(), rank = comm.rank();
    const int num_local = 2 * size;
    std::vector<int> ranks( size ), neighbors( size );
    std::vector<unsigned> ids( size );
    for ( int n = 0; n < size; ++n )
    {
        neighbors[n] = n;
        ranks[n] = n;
        ids[n] = 2 * n + 1;
    }
    DeviceArray<int> d_ranks( ranks );
    DeviceArray<unsigned> d_ids( ids );
    HaloT halo = use_topology ? HaloT( comm, num_local, UIntView( d_ids.p, size ),
                                       IntView( d_ranks.p, size ), neighbors )
                              : HaloT( comm, num_local, UIntView( d_ids.p, size ),
                                       IntView( d_ranks.p, size ) );
    EXPECT_EQ( (int)halo.numLocal(), num_local );
    EXPECT_EQ( (int)halo.numGhost(), size );
    EXPECT_EQ( halo.numNeighbor(), size );
    EXPECT_EQ( halo.neighborRank( 0 ), rank ); // self first
    const int n_tot = num_local + size;
    std::vector<int> hi( n_tot, 0 );
    std::vector<double> hd( 2 * n_tot, 0.0 );
    for ( int i = 0; i < num_local; ++i )
    {
        hi[i] = rank + 1;
        hd[2 * i] = rank + 1;
        hd[2 * i + 1] = rank + 1.5;
    }
    DeviceArray<int> d_i( hi );
    DeviceArray<double> d_d( hd );
    IntView s_int( d_i.p, n_tot );
    DblView2 s_dbl( d_d.p, n_tot );
    // gather both members as one tuple (the AoSoA gather), then check (tstHalo.hpp:183-222)
    Cabana::gather( halo, s_int, s_dbl );
    cudaStreamSynchronize( comm.stream() );
    auto gi = d_i.host();
    auto gd = d_d.host();
    for ( int i = 0; i < num_local; ++i )
        EXPECT_TRUE( gi[i] == rank + 1 && gd[2 * i] == rank + 1 && gd[2 * i + 1] == rank + 1.5 );
    for ( int i = num_local; i < n_tot; ++i )
    {
        // ghosts in neighbour order: with a topology getUniqueTopology swaps this rank with
        // rank 0's slot; without one the export ranks are ascending with self swapped to front
        // -- the same list here
        const int src = halo.neighborRank( i - num_local );
        const int send_rank = i - num_local;
        const int expect_src = send_rank == 0 ? rank : ( send_rank == rank ? 0 : send_rank );
        EXPECT_EQ( src, expect_src );
        EXPECT_TRUE( gi[i] == src + 1 && gd[2 * i] == src + 1 && gd[2 * i + 1] == src + 1.5 );
    }
    // scatter: every ghost had a unique destination, so ghosted elements double (:238-258)
    Cabana::scatter( halo, s_int );
    Cabana::scatter( halo, s_dbl );
    cudaStreamSynchronize( comm.stream() );
    gi = d_i.host();
    gd = d_d.host();
    for ( int i = 0; i < size; ++i )
    {
        EXPECT_TRUE( gi[2 * i] == rank + 1 && gd[4 * i] == rank + 1 );
        EXPECT_TRUE( gi[2 * i + 1] == 2 * ( rank + 1 ) && gd[2 * ( 2 * i + 1 )] == 2 * ( rank + 1 ) &&
                     gd[2 * ( 2 * i + 1 ) + 1] == 2 * ( rank + 1.5 ) );
    }
    // slice of the wrong size is refused (impl/Cabana_Halo_Mpi.hpp:57-59)
    bool threw = false;
    try
    {
        Cabana::gather( halo, IntView( d_i.p, n_tot - 1 ) );
    }
    catch ( const std::runtime_error& )
    {
        threw = true;
    }
    EXPECT_TRUE( threw );
}

// ---- tstHalo "all": every rank sends its single element to every rank -> scatter collisions ---
static void testHaloAll( const Cabana::NcclCommunicator& comm, bool use_topology )
{
    const int size = comm.size(), rank = comm.rank();
    std::vector<int> ranks( size ), neighbors( size );
    for ( int n = 0; n < size; ++n )
        neighbors[n] = ranks[n] = n;
    DeviceArray<int> d_ranks( ranks );
    DeviceArray<unsigned> d_ids( std::vector<unsigned>( size, 0u ) );
    HaloT halo = use_topology ? HaloT( comm, 1, UIntView( d_ids.p, size ),
                                       IntView( d_ranks.p, size ), neighbors )
                              : HaloT( comm, 1, UIntView( d_ids.p, size ),
                                       IntView( d_ranks.p, size ) );
    const int n_tot = 1 + size;
    std::vector<double> hd( 2 * n_tot, 0.0 );
    hd[0] = rank + 1;
    hd[1] = rank + 1.5;
    DeviceArray<double> d_d( hd );
    std::vector<float> hf( n_tot, 0.f );
    hf[0] = (float)( rank + 1 );
    DeviceArray<float> d_f( hf );
    DblView2 s_dbl( d_d.p, n_tot );
    Cabana::View2D<float, 1> s_flt( d_f.p, n_tot );
    Cabana::gather( halo, s_dbl );
    Cabana::gather( halo, s_flt );
    Cabana::scatter( halo, s_dbl );
    Cabana::scatter( halo, s_flt ); // typed scatter: float
    cudaStreamSynchronize( comm.stream() );
    auto gd = d_d.host();
    auto gf = d_f.host();
    // every rank (itself included) returned my value once: original + size * original
    EXPECT_EQ( gd[0], ( rank + 1 ) * ( size + 1.0 ) );
    EXPECT_EQ( gd[1], ( rank + 1.5 ) * ( size + 1.0 ) );
    EXPECT_EQ( gf[0], (float)( ( rank + 1 ) * ( size + 1 ) ) );
    for ( int n = 0; n < size; ++n )
        EXPECT_EQ( gd[2 * ( 1 + n )], halo.neighborRank( n ) + 1.0 );
}

// ---- tstDistributor test1: everything stays; test2: every other element stays (in place) ----
static void testDistributorStay( const Cabana::NcclCommunicator& comm, bool use_topology )
{
    const int rank = comm.rank();
    const int num_data = 10;
    std::vector<int> stay( num_data, rank ), half( num_data );
    for ( int n = 0; n < num_data; ++n )
        half[n] = ( n % 2 == 0 ) ? rank : -1;
    std::vector<int> nb( 1, rank );
    std::vector<int> hi( num_data );
    std::vector<double> hd( 2 * num_data );
    for ( int i = 0; i < num_data; ++i )
    {
        hi[i] = rank + i;
        hd[2 * i] = rank + i;
        hd[2 * i + 1] = rank + i + 0.5;
    }
    {
        DeviceArray<int> d_r( stay ), s_i( hi ), t_i( num_data );
        DeviceArray<double> s_d( hd ), t_d( 2 * num_data );
        DistT dist = use_topology ? DistT( comm, IntView( d_r.p, num_data ), nb )
                                  : DistT( comm, IntView( d_r.p, num_data ) );
        EXPECT_EQ( (int)dist.totalNumImport(), num_data );
        EXPECT_EQ( dist.numNeighbor(), 1 );
        Cabana::migrate( dist, std::make_tuple( IntView( s_i.p, num_data ), DblView2( s_d.p, num_data ) ),
                         std::make_tuple( IntView( t_i.p, num_data ), DblView2( t_d.p, num_data ) ) );
        cudaStreamSynchronize( comm.stream() );
        auto steering = std::vector<unsigned>( num_data );
        cudaMemcpy( steering.data(), dist.getExportSteering(), sizeof( unsigned ) * num_data,
                    cudaMemcpyDeviceToHost );
        auto gi = t_i.host();
        auto gd = t_d.host();
        for ( int i = 0; i < num_data; ++i )
            EXPECT_TRUE( gi[i] == rank + (int)steering[i] && gd[2 * i] == rank + (int)steering[i] &&
                         gd[2 * i + 1] == rank + (int)steering[i] + 0.5 );
    }
    {
        // test2, in place: 5 of 10 elements survive, the storage shrinks
        DeviceArray<int> d_r( half ), s_i( hi );
        DeviceArray<double> s_d( hd );
        DistT dist = use_topology ? DistT( comm, IntView( d_r.p, num_data ), nb )
                                  : DistT( comm, IntView( d_r.p, num_data ) );
        const std::size_t new_size = Cabana::migrate_in_place(
            dist, num_data, IntView( s_i.p, num_data ), DblView2( s_d.p, num_data ) );
        cudaStreamSynchronize( comm.stream() );
        EXPECT_EQ( (int)new_size, num_data / 2 );
        std::vector<unsigned> steering( new_size );
        cudaMemcpy( steering.data(), dist.getExportSteering(), sizeof( unsigned ) * new_size,
                    cudaMemcpyDeviceToHost );
        auto gi = s_i.host();
        auto gd = s_d.host();
        for ( std::size_t i = 0; i < new_size; ++i )
            EXPECT_TRUE( gi[i] == rank + (int)steering[i] && gd[2 * i + 1] == rank + (int)steering[i] + 0.5 );
    }
}

// ---- tstDistributor test4: two elements to every rank, interleaved; plus a ring -------------
static void testDistributorAllToAll( const Cabana::NcclCommunicator& comm, bool use_topology )
{
    const int size = comm.size(), rank = comm.rank();
    const int num_data = 2 * size;
    std::vector<int> ranks( num_data ), nb( size );
    for ( int n = 0; n < size; ++n )
    {
        ranks[n] = n;
        ranks[n + size] = n;
        nb[n] = n;
    }
    DeviceArray<int> d_r( ranks );
    DistT dist = use_topology ? DistT( comm, IntView( d_r.p, num_data ), nb )
                              : DistT( comm, IntView( d_r.p, num_data ) );
    EXPECT_EQ( (int)dist.totalNumImport(), num_data );
    std::vector<int> hi( num_data, rank );
    std::vector<double> hd( 2 * num_data );
    for ( int i = 0; i < num_data; ++i )
    {
        hd[2 * i] = rank;
        hd[2 * i + 1] = rank + 0.5;
    }
    DeviceArray<int> s_i( hi ), t_i( num_data );
    DeviceArray<double> s_d( hd ), t_d( 2 * num_data );
    Cabana::migrate( dist, IntView( s_i.p, num_data ), IntView( t_i.p, num_data ) );
    Cabana::migrate( dist, DblView2( s_d.p, num_data ), DblView2( t_d.p, num_data ) );
    cudaStreamSynchronize( comm.stream() );
    auto gi = t_i.host();
    auto gd = t_d.host();
    // two elements from every neighbour, blocks in neighbour order (self first)
    for ( int n = 0; n < dist.numNeighbor(); ++n )
        for ( int k = 0; k < 2; ++k )
        {
            const int i = 2 * n + k, src = dist.neighborRank( n );
            EXPECT_TRUE( gi[i] == src && gd[2 * i] == src && gd[2 * i + 1] == src + 0.5 );
        }
    EXPECT_EQ( dist.neighborRank( 0 ), rank );
    // ring: everything to the next rank (no self-send unless the world is one rank)
    const int n_ring = 100;
    std::vector<int> next( n_ring, ( rank + 1 ) % size );
    DeviceArray<int> d_next( next );
    DistT ring( comm, IntView( d_next.p, n_ring ) );
    EXPECT_EQ( (int)ring.totalNumImport(), n_ring );
    std::vector<double> v( 2 * n_ring, (double)rank );
    DeviceArray<double> s( v ), t( 2 * n_ring );
    Cabana::migrate( ring, DblView2( s.p, n_ring ), DblView2( t.p, n_ring ) );
    cudaStreamSynchronize( comm.stream() );
    auto got = t.host();
    bool ok = true;
    for ( auto x : got )
        ok = ok && x == (double)( ( rank + size - 1 ) % size );
    EXPECT_TRUE( ok );
}

// ---- SlabPeerHalo::step: ghosts through peer-memory windows + owner-local VerletList ----------
// Every rank draws the same pseudo-random particles, owns those of its x slab, and the union of
// the owner-local lists (ghost ids mapped back through a global-id member that travels with the
// positions) must be the brute-force neighbour list of the whole box.
static void testSlabPeerStep( const Cabana::NcclCommunicator& comm )
{
    const int size = comm.size(), rank = comm.rank();
    const double r = 3.0, slab = 12.0, L = slab * size, Lyz = 18.0;
    const int n_all = 3000 * size;
    std::vector<double> X( 3 * (std::size_t)n_all );
    unsigned long long sd = 12345ull;
    auto rnd = [&]()
    {
        sd = sd * 6364136223846793005ull + 1442695040888963407ull;
        return (double)( sd >> 11 ) / 9007199254740992.0;
    };
    for ( int i = 0; i < n_all; ++i )
    {
        X[3 * i] = rnd() * L;
        X[3 * i + 1] = rnd() * Lyz;
        X[3 * i + 2] = rnd() * Lyz;
    }
    std::vector<double> bounds( size + 1 );
    for ( int g = 0; g <= size; ++g )
        bounds[g] = slab * g;
    std::vector<int> own;
    for ( int i = 0; i < n_all; ++i )
    {
        int g = (int)( X[3 * i] / slab );
        g = g > size - 1 ? size - 1 : g;
        if ( g == rank )
            own.push_back( i );
    }
    const std::size_t nl = own.size(), cap_face = nl + 16, cap = nl + 2 * cap_face;
    std::vector<double> hx( 3 * cap, 0.0 );
    std::vector<int> hg( cap, -1 );
    for ( std::size_t k = 0; k < nl; ++k )
    {
        for ( int d = 0; d < 3; ++d )
            hx[3 * k + d] = X[3 * own[k] + d];
        hg[k] = own[k];
    }
    DeviceArray<double> dx( hx );
    DeviceArray<int> dg( hg );
    Cabana::View2D<double, 3> x_all( dx.p, cap );
    IntView gid( dg.p, cap );
    cb_field f[2] = { x_all.field(), gid.field() };
    const std::size_t tb = (std::size_t)cb_comm_tuple_bytes( f, 2 );
    Cabana::SlabPeerHalo halo( comm, bounds, r, cap_face, tb );
    using ListT = Cabana::VerletList<Cabana::DeviceSpace, Cabana::FullNeighborTag,
                                     Cabana::VerletLayoutCSR, Cabana::TeamOpTag>;
    ListT list;
    const double gmin[3] = { halo.localGridMinX(), 0.0, 0.0 };
    const double gmax[3] = { halo.localGridMaxX(), Lyz, Lyz };
    for ( int rep = 0; rep < 3; ++rep ) // (rebuilds take the speculative fill pass)
    {
        const auto ng = halo.step( list, x_all, nl, r, 1.0, gmin, gmax, gid );
        cudaStreamSynchronize( comm.stream() );
        const std::size_t ntot = nl + ng.first + ng.second;
        EXPECT_EQ( list._data.num_particles, ntot );
        if ( size > 1 )
            EXPECT_TRUE( ng.first + ng.second > 0 );
        else
            EXPECT_EQ( ng.first + ng.second, (std::size_t)0 );
        std::vector<int> counts( ntot ), offsets( ntot ), nb( list._data.total ? list._data.total : 1 );
        cudaMemcpy( counts.data(), list._data.counts, sizeof( int ) * ntot, cudaMemcpyDeviceToHost );
        cudaMemcpy( offsets.data(), list._data.offsets, sizeof( int ) * ntot, cudaMemcpyDeviceToHost );
        cudaMemcpy( nb.data(), list._data.neighbors, sizeof( int ) * list._data.total,
                    cudaMemcpyDeviceToHost );
        const auto g_all = dg.host();
        int bad = 0;
        for ( std::size_t k = nl; k < ntot; ++k )
            bad += counts[k] != 0; // ghost rows stay empty
        for ( std::size_t k = 0; k < nl && bad == 0; ++k )
        {
            const int i = own[k];
            std::vector<int> want, got;
            for ( int j = 0; j < n_all; ++j )
            {
                if ( j == i )
                    continue;
                const double ddx = X[3 * i] - X[3 * j], ddy = X[3 * i + 1] - X[3 * j + 1],
                             ddz = X[3 * i + 2] - X[3 * j + 2];
                volatile double sx = ddx * ddx, sy = ddy * ddy, sz = ddz * ddz; // (no contraction)
                if ( ( sx + sy ) + sz <= r * r )
                    want.push_back( j );
            }
            for ( int c = 0; c < counts[k]; ++c )
                got.push_back( g_all[nb[offsets[k] + c]] );
            std::sort( got.begin(), got.end() );
            if ( got != want )
                ++bad;
        }
        EXPECT_EQ( bad, 0 );
    }
    halo.close();
}

int main( int argc, char** argv )
{
    if ( argc < 4 )
    {
        std::printf( "usage: %s <rank> <world> <id-file>\n", argv[0] );
        return 2;
    }
    const int rank = std::atoi( argv[1] ), world = std::atoi( argv[2] );
    g_rank = rank;
    int ndev = 0;
    cudaGetDeviceCount( &ndev );
    if ( ndev < world )
    {
        std::printf( "need %d GPUs, have %d\n", world, ndev );
        return 3;
    }
    cudaSetDevice( rank );
    cb_set_device( rank );
    ncclUniqueId id;
    if ( rank == 0 )
    {
        ncclGetUniqueId( &id );
        std::ofstream f( std::string( argv[3] ) + ".tmp", std::ios::binary );
        f.write( reinterpret_cast<const char*>( &id ), sizeof( id ) );
        f.close();
        std::rename( ( std::string( argv[3] ) + ".tmp" ).c_str(), argv[3] );
    }
    else
    {
        for ( int tries = 0;; ++tries )
        {
            std::ifstream f( argv[3], std::ios::binary );
            if ( f && f.read( reinterpret_cast<char*>( &id ), sizeof( id ) ) )
                break;
            if ( tries > 600 )
            {
                std::printf( "rank %d: no id file\n", rank );
                return 4;
            }
            std::this_thread::sleep_for( std::chrono::milliseconds( 100 ) );
        }
    }
    ncclComm_t nc;
    if ( ncclCommInitRank( &nc, world, id, rank ) != ncclSuccess )
    {
        std::printf( "rank %d: ncclCommInitRank failed\n", rank );
        return 5;
    }
    cudaStream_t st;
    cudaStreamCreateWithFlags( &st, cudaStreamNonBlocking );
    try
    {
        Cabana::NcclCommunicator comm( nc, st );
        for ( int topo = 0; topo < 2; ++topo )
        {
            testHaloUnique( comm, topo != 0 );
            testHaloAll( comm, topo != 0 );
            testDistributorStay( comm, topo != 0 );
            testDistributorAllToAll( comm, topo != 0 );
        }
        testSlabPeerStep( comm );
    }
    catch ( const std::exception& e )
    {
        std::printf( "FAIL rank %d: exception %s\n", rank, e.what() );
        ++g_fail;
    }
    cudaStreamSynchronize( st );
    ncclCommDestroy( nc );
    cudaStreamDestroy( st );
    if ( g_fail == 0 )
        std::printf( "rank %d/%d: ALL CABANA COMM TESTS PASSED\n", rank, world );
    return g_fail == 0 ? 0 : 1;
}
