// =============================================================================
// Cabana_B200_Comm.hpp -- Cabana::Halo / Cabana::Distributor / gather / scatter / migrate
// for the B200-native library, templated on a `Nccl` communication-space tag.
//
// Mirrors, for the hot path's exchange step, the reference's
//   CommunicationPlan<MemorySpace, CommSpace>   core/src/Cabana_CommunicationPlanBase.hpp:396-700,
//                                               core/src/impl/Cabana_CommunicationPlan_Mpi.hpp:60-410
//   Distributor<MemorySpace, CommSpace>         core/src/Cabana_Distributor.hpp:62-146
//   migrate( distributor, src, dst ) / in place core/src/Cabana_Distributor.hpp:275-414,
//                                               core/src/impl/Cabana_Migrate_Mpi.hpp:41-177
//   Halo<MemorySpace, Export, CommSpace>        core/src/Cabana_Halo.hpp:59-172
//   gather( halo, slices ) / scatter( halo, s ) core/src/Cabana_Halo.hpp:392-870,
//                                               core/src/impl/Cabana_Halo_Mpi.hpp:41-350
//   comm-space tags                             core/src/Cabana_Tags.hpp:21-47
// with MPI replaced by NCCL over NVLink: the per-neighbour MPI_Send/Irecv of one count becomes
// ONE ncclAllGather of the export-count vector, the payload MPI_Isend/Irecv pairs become a
// grouped ncclSend/ncclRecv on the plan's stream, and pack / unpack / scatter-add are the
// library's CUDA kernels (cb_comm_* in cabana_b200.h).  No host staging, no Python.
//
// One process per GPU; the application creates the ncclComm_t (ncclCommInitRank with an id it
// distributes itself, or ncclCommInitAll) and hands it over in an NcclCommunicator.
// =============================================================================
#ifndef CABANA_B200_COMM_HPP
#define CABANA_B200_COMM_HPP

#include <algorithm>
#include <cstdint>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>

#include "Cabana_B200.hpp"

namespace Cabana
{

//! Communication-space tag: NCCL point-to-point over NVLink / NVSwitch (the reference's tags
//! are Mpi / MpiAdvance / StreamTriggered, Cabana_Tags.hpp:27-47).
struct Nccl
{
};
//! Plan construction tags (Cabana_Tags.hpp:21-25)
struct Export
{
};
struct Import
{
};

//! What MPI_Comm is to the reference: communicator + the stream its operations are ordered on.
class NcclCommunicator
{
  public:
    NcclCommunicator() = default;
    NcclCommunicator( ncclComm_t comm, cudaStream_t stream = nullptr )
        : _comm( comm )
        , _stream( stream )
    {
        if ( ncclCommUserRank( comm, &_rank ) != ncclSuccess ||
             ncclCommCount( comm, &_size ) != ncclSuccess )
            throw std::runtime_error( "Cabana::NcclCommunicator: invalid communicator" );
    }
    ncclComm_t comm() const { return _comm; }
    cudaStream_t stream() const { return _stream; }
    int rank() const { return _rank; }
    int size() const { return _size; }

  private:
    ncclComm_t _comm = nullptr;
    cudaStream_t _stream = nullptr;
    int _rank = 0, _size = 1;
};

namespace Impl
{
inline void ncclCheck( ncclResult_t r, const char* where )
{
    if ( r != ncclSuccess )
        throw std::runtime_error( std::string( where ) + ": " + ncclGetErrorString( r ) );
}
inline void cudaCheck( cudaError_t e, const char* where )
{
    if ( e != cudaSuccess )
        throw std::runtime_error( std::string( where ) + ": " + cudaGetErrorString( e ) );
}
//! Grow-only device buffer (CommunicationData's send/receive buffers,
//! Cabana_CommunicationPlanBase.hpp:700-960, persistent across calls).
struct DeviceBytes
{
    std::shared_ptr<char> ptr;
    std::size_t capacity = 0;
    char* ensure( std::size_t bytes )
    {
        if ( bytes > capacity )
        {
            ptr = device_alloc<char>( bytes );
            capacity = bytes;
        }
        return ptr.get();
    }
};
//! getUniqueTopology (Cabana_CommunicationPlanBase.hpp:374-394): sorted, unique, -1 removed,
//! this rank swapped with the first entry.
inline std::vector<int> getUniqueTopology( int my_rank, std::vector<int> topology )
{
    auto remove_end = std::remove( topology.begin(), topology.end(), -1 );
    std::sort( topology.begin(), remove_end );
    auto unique_end = std::unique( topology.begin(), remove_end );
    topology.resize( std::distance( topology.begin(), unique_end ) );
    for ( auto& n : topology )
        if ( n == my_rank )
        {
            std::swap( n, topology[0] );
            break;
        }
    return topology;
}
} // namespace Impl

//---------------------------------------------------------------------------//
// CommunicationPlan<MemorySpace, Nccl>
//---------------------------------------------------------------------------//
template <class MemorySpace, class CommSpace = Nccl>
class CommunicationPlan;

template <class MemorySpace>
class CommunicationPlan<MemorySpace, Nccl>
{
  public:
    using memory_space = MemorySpace;
    using size_type = unsigned int;

    explicit CommunicationPlan( const NcclCommunicator& comm )
        : _comm( comm )
    {
    }

    const NcclCommunicator& comm() const { return _comm; }
    int numNeighbor() const { return (int)_neighbors.size(); }
    int neighborRank( const int n ) const { return _neighbors[n]; }
    std::size_t numExport( const int n ) const { return _num_export[n]; }
    std::size_t totalNumExport() const { return _total_num_export; }
    std::size_t numImport( const int n ) const { return _num_import[n]; }
    std::size_t totalNumImport() const { return _total_num_import; }
    std::size_t exportSize() const { return _num_export_element; }
    //! Device pointer to the steering vector: export slot -> local element id, neighbour blocks in
    //! neighbour order (createExportSteering, :596-657).
    const std::uint32_t* getExportSteering() const { return _steering.get(); }
    //! First export slot of neighbour n in the steering vector / first import slot of neighbour n.
    std::size_t exportOffset( const int n ) const { return _export_offset[n]; }
    std::size_t importOffset( const int n ) const { return _import_offset[n]; }

    //! createFromExportsAndTopology (:452-470 -> impl :105-192) and createFromExportsOnly
    //! (:493-508 -> impl :270-410) in one: `neighbor_ranks` == nullptr means "no topology".
    //! element_export_ranks: DEVICE array of n_export ints (-1 = drop the element);
    //! element_export_ids: DEVICE array of the local ids to export (nullptr: element i itself).
    void create( const int* element_export_ranks, const std::uint32_t* element_export_ids,
                 const std::size_t num_export_element, const std::vector<int>* neighbor_ranks )
    {
        const int size = _comm.size(), rank = _comm.rank();
        cudaStream_t st = _comm.stream();
        _num_export_element = num_export_element;
        // countSendsAndCreateSteering (:96-224): counts per destination rank + the steering vector
        // grouped by ascending rank, deterministic order inside a block
        std::vector<int64_t> counts( size, 0 ), offsets( size + 1, 0 );
        auto steer_by_rank = Impl::device_alloc<std::uint32_t>( num_export_element );
        Impl::check( cb_comm_count_and_steer( element_export_ranks, (int64_t)num_export_element,
                                              size, counts.data(), offsets.data(),
                                              steer_by_rank.get(), element_export_ids, st ),
                     "Cabana::CommunicationPlan: countSendsAndCreateSteering" );
        // every rank learns every rank's export counts: ONE all-gather of `size` values
        std::vector<int64_t> matrix( (std::size_t)size * size, 0 );
        if ( size > 1 )
        {
            auto d_counts = Impl::device_alloc<int64_t>( (std::size_t)size * ( size + 1 ) );
            Impl::cudaCheck( cudaMemcpyAsync( d_counts.get(), counts.data(),
                                              sizeof( int64_t ) * size, cudaMemcpyHostToDevice,
                                              st ),
                             "Cabana::CommunicationPlan" );
            Impl::ncclCheck( ncclAllGather( d_counts.get(), d_counts.get() + size, size,
                                            ncclInt64, _comm.comm(), st ),
                             "Cabana::CommunicationPlan: ncclAllGather" );
            Impl::cudaCheck( cudaMemcpyAsync( matrix.data(), d_counts.get() + size,
                                              sizeof( int64_t ) * size * size,
                                              cudaMemcpyDeviceToHost, st ),
                             "Cabana::CommunicationPlan" );
            Impl::cudaCheck( cudaStreamSynchronize( st ), "Cabana::CommunicationPlan" );
        }
        else
            matrix[0] = counts[0];
        std::vector<int64_t> imports( size );
        for ( int s = 0; s < size; ++s )
            imports[s] = matrix[(std::size_t)s * size + rank]; // [source][destination]

        _neighbors.clear();
        if ( neighbor_ranks )
        {
            _neighbors = Impl::getUniqueTopology( rank, *neighbor_ranks );
            for ( int r = 0; r < size; ++r )
                if ( ( counts[r] > 0 || imports[r] > 0 ) &&
                     std::find( _neighbors.begin(), _neighbors.end(), r ) == _neighbors.end() )
                    throw std::logic_error(
                        "Cabana::CommunicationPlan: traffic with a rank outside the topology" );
        }
        else
        {
            for ( int r = 0; r < size; ++r )
                if ( counts[r] > 0 )
                    _neighbors.push_back( r );
            for ( auto& n : _neighbors ) // self-send first (impl :318-328)
                if ( n == rank )
                {
                    std::swap( n, _neighbors[0] );
                    break;
                }
            for ( int r = 0; r < size; ++r ) // import-only neighbours (impl :378-393)
                if ( imports[r] > 0 &&
                     std::find( _neighbors.begin(), _neighbors.end(), r ) == _neighbors.end() )
                    _neighbors.push_back( r );
        }
        const int num_n = (int)_neighbors.size();
        _num_export.assign( num_n, 0 );
        _num_import.assign( num_n, 0 );
        _export_offset.assign( num_n + 1, 0 );
        _import_offset.assign( num_n + 1, 0 );
        for ( int n = 0; n < num_n; ++n )
        {
            _num_export[n] = (std::size_t)counts[_neighbors[n]];
            _num_import[n] = (std::size_t)imports[_neighbors[n]];
            _export_offset[n + 1] = _export_offset[n] + _num_export[n];
            _import_offset[n + 1] = _import_offset[n] + _num_import[n];
        }
        _total_num_export = _export_offset[num_n];
        _total_num_import = _import_offset[num_n];
        // steering in NEIGHBOUR order (createExportSteering :626-637)
        _steering = Impl::device_alloc<std::uint32_t>( _total_num_export );
        for ( int n = 0; n < num_n; ++n )
            if ( _num_export[n] > 0 )
                Impl::check( cb_memcpy_d2d( _steering.get() + _export_offset[n],
                                            steer_by_rank.get() + offsets[_neighbors[n]],
                                            sizeof( std::uint32_t ) * _num_export[n], st ),
                             "Cabana::CommunicationPlan: steering" );
        Impl::cudaCheck( cudaStreamSynchronize( st ), "Cabana::CommunicationPlan" );
    }

    //! Exchange `tuple_bytes`-sized tuples: send block n (num_send[n] tuples at send + off) to
    //! neighbour n, receive block n from it; the self block is a device-to-device copy
    //! (impl/Cabana_Halo_Mpi.hpp:70-124: Irecv / Isend / Waitall -> one NCCL group).
    void exchange( const char* send, const std::vector<std::size_t>& send_offset,
                   const std::vector<std::size_t>& num_send, char* recv,
                   const std::vector<std::size_t>& recv_offset,
                   const std::vector<std::size_t>& num_recv, const std::size_t tuple_bytes ) const
    {
        cudaStream_t st = _comm.stream();
        const int num_n = numNeighbor();
        bool grouped = false;
        for ( int n = 0; n < num_n; ++n )
        {
            if ( _neighbors[n] == _comm.rank() )
            {
                if ( num_send[n] != num_recv[n] )
                    throw std::logic_error( "Cabana::CommunicationPlan: self block mismatch" );
                if ( num_send[n] > 0 )
                    Impl::check( cb_memcpy_d2d( recv + recv_offset[n] * tuple_bytes,
                                                send + send_offset[n] * tuple_bytes,
                                                num_send[n] * tuple_bytes, st ),
                                 "Cabana::CommunicationPlan: self copy" );
                continue;
            }
            if ( !grouped && ( num_send[n] > 0 || num_recv[n] > 0 ) )
            {
                Impl::ncclCheck( ncclGroupStart(), "ncclGroupStart" );
                grouped = true;
            }
            if ( num_recv[n] > 0 )
                Impl::ncclCheck( ncclRecv( recv + recv_offset[n] * tuple_bytes,
                                           num_recv[n] * tuple_bytes, ncclChar, _neighbors[n],
                                           _comm.comm(), st ),
                                 "ncclRecv" );
            if ( num_send[n] > 0 )
                Impl::ncclCheck( ncclSend( send + send_offset[n] * tuple_bytes,
                                           num_send[n] * tuple_bytes, ncclChar, _neighbors[n],
                                           _comm.comm(), st ),
                                 "ncclSend" );
        }
        if ( grouped )
            Impl::ncclCheck( ncclGroupEnd(), "ncclGroupEnd" );
    }

    //! Persistent buffers of the plan's gather/scatter/migrate calls.
    Impl::DeviceBytes& sendBuffer() const { return _send; }
    Impl::DeviceBytes& recvBuffer() const { return _recv; }

  protected:
    NcclCommunicator _comm;
    std::vector<int> _neighbors;
    std::vector<std::size_t> _num_export, _num_import, _export_offset, _import_offset;
    std::size_t _total_num_export = 0, _total_num_import = 0, _num_export_element = 0;
    std::shared_ptr<std::uint32_t> _steering;
    mutable Impl::DeviceBytes _send, _recv;
};

//---------------------------------------------------------------------------//
// Distributor (Cabana_Distributor.hpp:62-146)
//---------------------------------------------------------------------------//
template <class MemorySpace, class CommSpace = Nccl>
class Distributor : public CommunicationPlan<MemorySpace, CommSpace>
{
  public:
    //! With topology (:103-122).  element_export_ranks: View2D<int,1> / Slice<int,1> on the device
    //! holding the destination rank of every element (-1 drops it).
    template <class RankViewType>
    Distributor( const NcclCommunicator& comm, const RankViewType& element_export_ranks,
                 const std::vector<int>& neighbor_ranks )
        : CommunicationPlan<MemorySpace, CommSpace>( comm )
    {
        this->create( element_export_ranks.data(), nullptr, element_export_ranks.size(),
                      &neighbor_ranks );
    }
    //! Without topology (:124-146)
    template <class RankViewType>
    Distributor( const NcclCommunicator& comm, const RankViewType& element_export_ranks )
        : CommunicationPlan<MemorySpace, CommSpace>( comm )
    {
        this->create( element_export_ranks.data(), nullptr, element_export_ranks.size(), nullptr );
    }
};

//---------------------------------------------------------------------------//
// Halo, export-built (Cabana_Halo.hpp:59-172)
//---------------------------------------------------------------------------//
template <class MemorySpace, class BuildType = Export, class CommSpace = Nccl>
class Halo : public CommunicationPlan<MemorySpace, CommSpace>
{
    static_assert( std::is_same<BuildType, Export>::value,
                   "the C++ surface builds halos from exports; import-built halos are in the "
                   "Python mirror (cabana_b200.comm.Halo.from_imports)" );

  public:
    //! With topology (:107-129): element_export_ids / element_export_ranks are device arrays
    //! (View2D<unsigned,1> / View2D<int,1>) of equal size.
    template <class IdViewType, class RankViewType>
    Halo( const NcclCommunicator& comm, const std::size_t num_local,
          const IdViewType& element_export_ids, const RankViewType& element_export_ranks,
          const std::vector<int>& neighbor_ranks )
        : CommunicationPlan<MemorySpace, CommSpace>( comm )
        , _num_local( num_local )
    {
        if ( element_export_ids.size() != element_export_ranks.size() )
            throw std::runtime_error( "Cabana::Halo: Export ids and ranks different sizes!" );
        this->create( element_export_ranks.data(),
                      reinterpret_cast<const std::uint32_t*>( element_export_ids.data() ),
                      element_export_ranks.size(), &neighbor_ranks );
    }
    //! Without topology (:151-172)
    template <class IdViewType, class RankViewType>
    Halo( const NcclCommunicator& comm, const std::size_t num_local,
          const IdViewType& element_export_ids, const RankViewType& element_export_ranks )
        : CommunicationPlan<MemorySpace, CommSpace>( comm )
        , _num_local( num_local )
    {
        if ( element_export_ids.size() != element_export_ranks.size() )
            throw std::runtime_error( "Cabana::Halo: Export ids and ranks different sizes!" );
        this->create( element_export_ranks.data(),
                      reinterpret_cast<const std::uint32_t*>( element_export_ids.data() ),
                      element_export_ranks.size(), nullptr );
    }
    std::size_t numLocal() const { return _num_local; }
    std::size_t numGhost() const { return this->totalNumImport(); }

  private:
    std::size_t _num_local;
};

namespace Impl
{
template <class... Slices>
std::vector<cb_field> fields_of( const Slices&... s )
{
    return std::vector<cb_field>{ s.field()... };
}
inline std::vector<cb_field> with_size( std::vector<cb_field> f, std::size_t n )
{
    for ( auto& x : f )
        x.n = (int64_t)n;
    return f;
}
template <class T>
struct scatter_dtype;
template <>
struct scatter_dtype<double>
{
    static constexpr int value = CB_DTYPE_F64;
};
template <>
struct scatter_dtype<float>
{
    static constexpr int value = CB_DTYPE_F32;
};
template <>
struct scatter_dtype<int>
{
    static constexpr int value = CB_DTYPE_I32;
};
template <>
struct scatter_dtype<long long>
{
    static constexpr int value = CB_DTYPE_I64;
};
} // namespace Impl

//---------------------------------------------------------------------------//
// gather( halo, slices... ) (Cabana_Halo.hpp:677-682; impl/Cabana_Halo_Mpi.hpp:41-125): the
// slices hold numLocal() + numGhost() elements; ghosts land behind the local elements grouped
// by source rank in neighbour order.  Several slices travel as one tuple per element (what the
// reference does for an AoSoA).
//---------------------------------------------------------------------------//
template <class HaloType, class... Slices>
void gather( const HaloType& halo, const Slices&... slices )
{
    Impl::ScopedRegion region( "Cabana::gather" ); // impl/Cabana_Halo_Mpi.hpp:48
    auto fields = Impl::fields_of( slices... );
    for ( const auto& f : fields )
        if ( (std::size_t)f.n != halo.numLocal() + halo.numGhost() )
            throw std::runtime_error( "Cabana::gather: Slice is the wrong size" ); // :57-59
    const int nf = (int)fields.size();
    const std::size_t tb = (std::size_t)cb_comm_tuple_bytes( fields.data(), nf );
    cudaStream_t st = halo.comm().stream();
    char* send = halo.sendBuffer().ensure( halo.totalNumExport() * tb );
    char* recv = halo.recvBuffer().ensure( halo.totalNumImport() * tb );
    if ( halo.totalNumExport() > 0 )
        Impl::check( cb_comm_pack( fields.data(), nf, halo.getExportSteering(),
                                   (int64_t)halo.totalNumExport(), send, st ),
                     "Cabana::gather: pack" );
    const int num_n = halo.numNeighbor();
    std::vector<std::size_t> so( num_n ), ns( num_n ), ro( num_n ), nr( num_n );
    for ( int n = 0; n < num_n; ++n )
    {
        so[n] = halo.exportOffset( n );
        ns[n] = halo.numExport( n );
        ro[n] = halo.importOffset( n );
        nr[n] = halo.numImport( n );
    }
    halo.exchange( send, so, ns, recv, ro, nr, tb );
    if ( halo.totalNumImport() > 0 )
        Impl::check( cb_comm_unpack( fields.data(), nf, (int64_t)halo.numLocal(),
                                     (int64_t)halo.totalNumImport(), recv, st ),
                     "Cabana::gather: unpack" );
}

//---------------------------------------------------------------------------//
// scatter( halo, slice ) (Cabana_Halo.hpp:856-870; impl/Cabana_Halo_Mpi.hpp:236-350): ghost
// values go back to their owners and are atomically summed into them.
//---------------------------------------------------------------------------//
template <class HaloType, class SliceType>
void scatter( const HaloType& halo, const SliceType& slice )
{
    Impl::ScopedRegion region( "Cabana::scatter" ); // impl/Cabana_Halo_Mpi.hpp:251
    cb_field f = slice.field();
    if ( (std::size_t)f.n != halo.numLocal() + halo.numGhost() )
        throw std::runtime_error( "Cabana::scatter: Slice is the wrong size" );
    const std::size_t tb = (std::size_t)cb_comm_tuple_bytes( &f, 1 ); // packed single-field tuples
    cudaStream_t st = halo.comm().stream();
    // send = the ghosts (import side of the plan), receive = one value per export slot
    char* send = halo.sendBuffer().ensure( halo.totalNumImport() * tb );
    char* recv = halo.recvBuffer().ensure( halo.totalNumExport() * tb );
    if ( halo.totalNumImport() > 0 )
        Impl::check( cb_comm_pack_range( &f, 1, (int64_t)halo.numLocal(),
                                         (int64_t)halo.totalNumImport(), send, st ),
                     "Cabana::scatter: pack" );
    const int num_n = halo.numNeighbor();
    std::vector<std::size_t> so( num_n ), ns( num_n ), ro( num_n ), nr( num_n );
    for ( int n = 0; n < num_n; ++n )
    {
        so[n] = halo.importOffset( n );
        ns[n] = halo.numImport( n );
        ro[n] = halo.exportOffset( n );
        nr[n] = halo.numExport( n );
    }
    halo.exchange( send, so, ns, recv, ro, nr, tb );
    if ( halo.totalNumExport() > 0 )
        Impl::check( cb_comm_scatter_add_typed(
                         &f, halo.getExportSteering(), (int64_t)halo.totalNumExport(), recv,
                         Impl::scatter_dtype<typename std::remove_const<
                             typename SliceType::value_type>::type>::value,
                         st ),
                     "Cabana::scatter: scatter_add" );
}

namespace Impl
{
//! migrate of one tuple set: pack by steering -> exchange -> unpack at [0, totalNumImport) of
//! the destination (staying elements first: this rank is neighbour 0 when it keeps anything,
//! impl/Cabana_Migrate_Mpi.hpp:92-172).
template <class DistributorType>
void migrate_fields( const DistributorType& d, std::vector<cb_field> src,
                     std::vector<cb_field> dst )
{
    ScopedRegion region( "Cabana::migrate" ); // impl/Cabana_Migrate_Mpi.hpp:48
    const int nf = (int)src.size();
    const std::size_t tb = (std::size_t)cb_comm_tuple_bytes( src.data(), nf );
    cudaStream_t st = d.comm().stream();
    char* send = d.sendBuffer().ensure( d.totalNumExport() * tb );
    char* recv = d.recvBuffer().ensure( d.totalNumImport() * tb );
    if ( d.totalNumExport() > 0 )
        check( cb_comm_pack( src.data(), nf, d.getExportSteering(), (int64_t)d.totalNumExport(),
                             send, st ),
               "Cabana::migrate: pack" );
    const int num_n = d.numNeighbor();
    std::vector<std::size_t> so( num_n ), ns( num_n ), ro( num_n ), nr( num_n );
    for ( int n = 0; n < num_n; ++n )
    {
        so[n] = d.exportOffset( n );
        ns[n] = d.numExport( n );
        ro[n] = d.importOffset( n );
        nr[n] = d.numImport( n );
    }
    d.exchange( send, so, ns, recv, ro, nr, tb );
    if ( d.totalNumImport() > 0 )
        check( cb_comm_unpack( dst.data(), nf, 0, (int64_t)d.totalNumImport(), recv, st ),
               "Cabana::migrate: unpack" );
}
} // namespace Impl

//---------------------------------------------------------------------------//
// migrate( distributor, src, dst ) for one slice (Cabana_Distributor.hpp:360-414) and for the
// member slices of an AoSoA moved together (:330-337): src holds exportSize() elements, dst
// totalNumImport().  Pass the members as std::tie-like packs: migrate( d, src0, dst0 ) or
// migrate( d, std::make_tuple( s0, s1 ), std::make_tuple( d0, d1 ) ).
//---------------------------------------------------------------------------//
template <class DistributorType, class SliceType>
auto migrate( const DistributorType& distributor, const SliceType& src, const SliceType& dst )
    -> decltype( src.field(), void() )
{
    if ( src.size() != distributor.exportSize() || dst.size() != distributor.totalNumImport() )
        throw std::runtime_error( "Cabana::migrate: Slice is the wrong size" ); // :374-377
    Impl::migrate_fields( distributor, Impl::fields_of( src ), Impl::fields_of( dst ) );
}
template <class DistributorType, class... S, class... D>
void migrate( const DistributorType& distributor, const std::tuple<S...>& src,
              const std::tuple<D...>& dst )
{
    auto fs = std::apply( []( const auto&... s ) { return Impl::fields_of( s... ); }, src );
    auto fd = std::apply( []( const auto&... s ) { return Impl::fields_of( s... ); }, dst );
    for ( const auto& f : fs )
        if ( (std::size_t)f.n != distributor.exportSize() )
            throw std::runtime_error( "Cabana::migrate: Source is the wrong size" );
    for ( const auto& f : fd )
        if ( (std::size_t)f.n != distributor.totalNumImport() )
            throw std::runtime_error( "Cabana::migrate: Destination is the wrong size" );
    Impl::migrate_fields( distributor, fs, fd );
}

//---------------------------------------------------------------------------//
// In-place migrate (Cabana_Distributor.hpp:275-305): the reference copies the AoSoA, resizes it
// to totalNumImport() and migrates the copy into it.  The member slices here are views of
// storage the caller owns, so the storage must hold max( exportSize(), totalNumImport() )
// elements (`capacity`); the call returns the new size.  The exported tuples are packed into
// the plan's send buffer BEFORE anything is unpacked, so no copy of the members is needed.
//---------------------------------------------------------------------------//
template <class DistributorType, class... Slices>
std::size_t migrate_in_place( const DistributorType& distributor, const std::size_t capacity,
                              const Slices&... members )
{
    if ( capacity < distributor.exportSize() || capacity < distributor.totalNumImport() )
        throw std::runtime_error( "Cabana::migrate: in-place storage is too small" );
    auto f = Impl::fields_of( members... );
    Impl::migrate_fields( distributor, Impl::with_size( f, distributor.exportSize() ),
                          Impl::with_size( f, distributor.totalNumImport() ) );
    return distributor.totalNumImport();
}

//---------------------------------------------------------------------------//
// SlabPeerHalo: the sharded step of a 1-D x-slab decomposition from host C++.  Replaces, for the
// slab topology, Halo construction + gather( halo, x ) + VerletList::build( x, 0, num_local, ... )
// (Cabana_Halo.hpp:107-115, impl/Cabana_Halo_Mpi.hpp:41-125, Cabana_VerletList.hpp:1351-1392)
// by ONE call: ghosts are packed straight into the neighbours' HBM windows (CUDA IPC over
// NVLink), waited for and unpacked on the device, and the owner-local list is built with the
// ghost counts never visiting the host (cb_slab_step).  NCCL is used once, to exchange the
// window handles.  One process per GPU of one node.
//---------------------------------------------------------------------------//
class SlabPeerHalo
{
  public:
    //! bounds: size()+1 ascending slab faces; halo_width: the cutoff the ghosts must cover;
    //! capacity_tuples: room per face; tuple_bytes: cb_comm_tuple_bytes of the slices that travel.
    SlabPeerHalo( const NcclCommunicator& comm, const std::vector<double>& bounds,
                  const double halo_width, const std::size_t capacity_tuples,
                  const std::size_t tuple_bytes )
        : _comm( comm )
        , _capacity( capacity_tuples )
        , _tuple_bytes( tuple_bytes )
    {
        const int size = comm.size(), rank = comm.rank();
        if ( (int)bounds.size() != size + 1 )
            throw std::runtime_error( "Cabana::SlabPeerHalo: need size + 1 slab faces" );
        _lo = bounds[rank];
        _hi = bounds[rank + 1];
        _hw = halo_width * ( 1.0 + 0x1p-40 );
        _lo_rank = rank > 0 ? rank - 1 : -1;
        _hi_rank = rank < size - 1 ? rank + 1 : -1;
        for ( int g = 0; g < size; ++g )
            if ( size > 1 && bounds[g + 1] - bounds[g] < _hw )
                throw std::runtime_error( "Cabana::SlabPeerHalo: a slab is thinner than the halo" );
        // A pusher addresses the neighbour's window with ITS OWN idea of the layout, so every
        // rank must use the same capacity: the largest one asked for.
        if ( size > 1 )
        {
            auto d = Impl::device_alloc<long long>( 1 );
            const long long mine_cap = (long long)_capacity;
            long long agreed = mine_cap;
            cudaStream_t st = comm.stream();
            Impl::cudaCheck( cudaMemcpyAsync( d.get(), &mine_cap, sizeof( long long ),
                                              cudaMemcpyHostToDevice, st ),
                             "SlabPeerHalo: capacity" );
            Impl::ncclCheck( ncclAllReduce( d.get(), d.get(), 1, ncclInt64, ncclMax, comm.comm(), st ),
                             "SlabPeerHalo: capacity" );
            Impl::cudaCheck( cudaMemcpyAsync( &agreed, d.get(), sizeof( long long ),
                                              cudaMemcpyDeviceToHost, st ),
                             "SlabPeerHalo: capacity" );
            Impl::cudaCheck( cudaStreamSynchronize( st ), "SlabPeerHalo: capacity" );
            _capacity = (std::size_t)agreed;
        }
        // my receive windows; handles travel as 2 x 64 bytes per rank: [from_lo | from_hi]
        unsigned char mine[2 * CB_IPC_HANDLE_BYTES] = { 0 };
        if ( _lo_rank >= 0 )
        {
            Impl::check( cb_p2p_window_create( &_from_lo, (int64_t)_capacity, (int64_t)_tuple_bytes ),
                         "Cabana::SlabPeerHalo: window" );
            Impl::check( cb_p2p_window_get_handle( _from_lo, mine ), "cb_p2p_window_get_handle" );
        }
        if ( _hi_rank >= 0 )
        {
            Impl::check( cb_p2p_window_create( &_from_hi, (int64_t)_capacity, (int64_t)_tuple_bytes ),
                         "Cabana::SlabPeerHalo: window" );
            Impl::check( cb_p2p_window_get_handle( _from_hi, mine + CB_IPC_HANDLE_BYTES ),
                         "cb_p2p_window_get_handle" );
        }
        std::vector<unsigned char> all( (std::size_t)size * sizeof( mine ) );
        if ( size > 1 )
        {
            auto d_send = Impl::device_alloc<unsigned char>( sizeof( mine ) );
            auto d_recv = Impl::device_alloc<unsigned char>( all.size() );
            cudaStream_t st = comm.stream();
            Impl::cudaCheck( cudaMemcpyAsync( d_send.get(), mine, sizeof( mine ),
                                              cudaMemcpyHostToDevice, st ),
                             "SlabPeerHalo: handle upload" );
            Impl::ncclCheck( ncclAllGather( d_send.get(), d_recv.get(), sizeof( mine ), ncclChar,
                                            comm.comm(), st ),
                             "SlabPeerHalo: ncclAllGather" );
            Impl::cudaCheck( cudaMemcpyAsync( all.data(), d_recv.get(), all.size(),
                                              cudaMemcpyDeviceToHost, st ),
                             "SlabPeerHalo: handle download" );
            Impl::cudaCheck( cudaStreamSynchronize( st ), "SlabPeerHalo: sync" );
            // my low-face layer lands in the lower neighbour's "from_hi" window and vice versa
            if ( _lo_rank >= 0 )
                Impl::check( cb_p2p_window_open( all.data() + (std::size_t)_lo_rank * sizeof( mine ) +
                                                     CB_IPC_HANDLE_BYTES,
                                                 &_peer_lo ),
                             "Cabana::SlabPeerHalo: cannot map the lower neighbour's window" );
            if ( _hi_rank >= 0 )
                Impl::check( cb_p2p_window_open( all.data() + (std::size_t)_hi_rank * sizeof( mine ),
                                                 &_peer_hi ),
                             "Cabana::SlabPeerHalo: cannot map the upper neighbour's window" );
            barrier(); // every window is mapped before anyone pushes
        }
    }
    SlabPeerHalo( const SlabPeerHalo& ) = delete;
    SlabPeerHalo& operator=( const SlabPeerHalo& ) = delete;
    //! Collective: no neighbour may still be pushing when the windows go away.
    void close()
    {
        if ( _closed )
            return;
        cudaStreamSynchronize( _comm.stream() );
        if ( _comm.size() > 1 )
            barrier();
        if ( _peer_lo )
            cb_p2p_window_close( _peer_lo );
        if ( _peer_hi )
            cb_p2p_window_close( _peer_hi );
        if ( _from_lo )
            cb_p2p_window_destroy( _from_lo );
        if ( _from_hi )
            cb_p2p_window_destroy( _from_hi );
        _peer_lo = _peer_hi = nullptr;
        _from_lo = _from_hi = nullptr;
        _closed = true;
    }
    ~SlabPeerHalo()
    {
        // (not collective: call close() on every rank before the object dies)
        if ( !_closed )
        {
            cudaStreamSynchronize( _comm.stream() );
            if ( _peer_lo )
                cb_p2p_window_close( _peer_lo );
            if ( _peer_hi )
                cb_p2p_window_close( _peer_hi );
            if ( _from_lo )
                cb_p2p_window_destroy( _from_lo );
            if ( _from_hi )
                cb_p2p_window_destroy( _from_hi );
        }
    }

    //! The local grid of this rank along x: the slab widened by the halo on its interior faces.
    double localGridMinX() const { return _lo_rank >= 0 ? _lo - _hw : _lo; }
    double localGridMaxX() const { return _hi_rank >= 0 ? _hi + _hw : _hi; }

    //! One step: gather the ghosts of `members` (the first must be the positions `x_all`; all hold
    //! room for owned + ghost elements) and build `list` over x_all[0 : num_local + ghosts) with
    //! begin = 0, end = num_local.  Returns { ghosts from the lower, from the upper neighbour }.
    template <class VerletListType, class PositionSlice, class ArrayType, class... Slices>
    std::pair<std::size_t, std::size_t>
    step( VerletListType& list, const PositionSlice& x_all, const std::size_t num_local,
          const double neighborhood_radius, const double cell_size_ratio,
          const ArrayType& grid_min, const ArrayType& grid_max, const Slices&... more_members )
    {
        auto fields = Impl::fields_of( x_all, more_members... );
        const int nf = (int)fields.size();
        if ( (std::size_t)cb_comm_tuple_bytes( fields.data(), nf ) != _tuple_bytes )
            throw std::runtime_error( "Cabana::SlabPeerHalo::step: members do not match the windows" );
        const std::size_t need = 2 * ( num_local ? num_local : 1 );
        if ( _steer_n < need )
        {
            _steer = Impl::device_alloc<std::uint32_t>( need + need / 8 );
            _steer_n = need + need / 8;
        }
        ++_seq;
        const cb_positions xd = x_all.positions();
        const double mn[3] = { (double)grid_min[0], (double)grid_min[1], (double)grid_min[2] };
        const double mx[3] = { (double)grid_max[0], (double)grid_max[1], (double)grid_max[2] };
        int64_t counts[2] = { 0, 0 };
        list.buildThrough(
            [&]( cb_verlet* h, int algorithm, int layout, int build_op )
            {
                Impl::check( cb_slab_step( h, &xd, fields.data(), nf, (int64_t)num_local, _lo + _hw,
                                           _hi - _hw, _peer_lo, _peer_hi, _from_lo, _from_hi,
                                           (int64_t)_capacity, _seq, _steer.get(),
                                           neighborhood_radius, cell_size_ratio, mn, mx, 0,
                                           algorithm, layout, build_op, counts, _comm.stream() ),
                             "Cabana::SlabPeerHalo::step" );
            } );
        return { (std::size_t)counts[0], (std::size_t)counts[1] };
    }

  private:
    void barrier()
    {
        auto d = Impl::device_alloc<int>( 1 );
        cudaStream_t st = _comm.stream();
        Impl::cudaCheck( cudaMemsetAsync( d.get(), 0, sizeof( int ), st ), "SlabPeerHalo: barrier" );
        Impl::ncclCheck( ncclAllReduce( d.get(), d.get(), 1, ncclInt, ncclSum, _comm.comm(), st ),
                         "SlabPeerHalo: barrier" );
        Impl::cudaCheck( cudaStreamSynchronize( st ), "SlabPeerHalo: barrier" );
    }
    NcclCommunicator _comm;
    std::size_t _capacity, _tuple_bytes;
    double _lo = 0.0, _hi = 0.0, _hw = 0.0;
    int _lo_rank = -1, _hi_rank = -1;
    cb_p2p_window* _from_lo = nullptr;
    cb_p2p_window* _from_hi = nullptr;
    void* _peer_lo = nullptr;
    void* _peer_hi = nullptr;
    std::shared_ptr<std::uint32_t> _steer;
    std::size_t _steer_n = 0;
    std::uint64_t _seq = 0;
    bool _closed = false;
};

} // namespace Cabana

#endif // CABANA_B200_COMM_HPP
