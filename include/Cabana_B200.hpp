// =============================================================================
// Cabana_B200.hpp -- header-only C++17 shim that reproduces Cabana's C++ API
// surface for the neighbour-list hot path on top of the C ABI (cabana_b200.h).
//
// Cabana itself is C++ templates over Kokkos; this header mirrors, with the same
// names, template parameters, argument order and meaning:
//   LinkedCellList<MemorySpace,Scalar,3>, createLinkedCellList, permute
//                                    core/src/Cabana_LinkedCellList.hpp:128-1145
//   VerletList<MemorySpace,AlgorithmTag,LayoutTag,BuildTag,3>, createVerletList,
//   VerletListData<..>, VerletLayoutCSR/2D
//                                    core/src/Cabana_VerletList.hpp:30-113, :824-1597
//   FullNeighborTag / HalfNeighborTag, NeighborList<ListType> traits
//                                    core/src/Cabana_NeighborList.hpp:35-252,
//                                    core/src/Cabana_VerletList.hpp:1603-1698
//   FirstNeighborsTag, SerialOpTag, TeamOpTag, TeamVectorOpTag,
//   neighbor_parallel_for / neighbor_parallel_reduce / for_each_neighbor
//                                    core/src/Cabana_Parallel.hpp:190-212, :251-293,
//                                    :386-435, :638-685, :787-844, :1058-1072
// so that code written against Cabana's path compiles against this header with the
// Kokkos pieces replaced by the small stand-ins below (DeviceSpace, RangePolicy,
// Slice, View2D).  All device work happens in libcabana_b200.so (hand-written
// sm_100a kernels); generic user functors are instantiated here as CUDA kernels
// when this header is compiled by nvcc (a functor cannot cross a C ABI).
//
// There is no CPU fallback: every build/traversal call goes to the GPU library and
// throws std::runtime_error with cb_last_error_string() if it fails.
// =============================================================================
#ifndef CABANA_B200_HPP
#define CABANA_B200_HPP

#include <algorithm>
#include <array>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "cabana_b200.h"

#if defined( __has_include )
#if __has_include( <nvtx3/nvToolsExt.h> )
#include <nvtx3/nvToolsExt.h>
#define CABANA_B200_HAVE_NVTX 1
#endif
#endif

#if defined( __CUDACC__ )
#include <cuda_runtime.h>
#define CABANA_B200_FUNCTION __host__ __device__
#define CABANA_B200_DEVICE __device__
#else
#define CABANA_B200_FUNCTION
#define CABANA_B200_DEVICE
#endif

namespace Cabana
{

//---------------------------------------------------------------------------//
// Stand-ins for the Kokkos types on the path.
//---------------------------------------------------------------------------//
//! Device memory space (stands in for Kokkos::CudaSpace).
struct DeviceSpace
{
    using size_type = unsigned int; // Kokkos::CudaSpace::size_type
};

//! Execution space instance (stands in for Kokkos::Cuda): a CUDA stream.  Default-constructed it
//! is the default stream, like Kokkos::Cuda().  Everything that takes an ExecutionSpace in the
//! reference (VerletList constructors/build :1260-1281, :1351-1392; RangePolicy( space, b, e ))
//! launches its kernels on this stream.
class DeviceExecutionSpace
{
  public:
    DeviceExecutionSpace() = default;
    explicit DeviceExecutionSpace( cb_stream_t stream )
        : _stream( stream )
    {
    }
    cb_stream_t stream() const { return _stream; }
    //! Kokkos::Cuda::fence()
    void fence() const;

  private:
    cb_stream_t _stream = nullptr;
};

//! Kokkos::RangePolicy<ExecSpace[, WorkTag]> stand-in: [begin, end).
template <class WorkTag = void>
class RangePolicy
{
  public:
    using work_tag = WorkTag;
    using index_type = int;
    using execution_space = DeviceExecutionSpace;
    RangePolicy( std::size_t b, std::size_t e )
        : _b( b )
        , _e( e )
    {
    }
    RangePolicy( const DeviceExecutionSpace& space, std::size_t b, std::size_t e )
        : _space( space )
        , _b( b )
        , _e( e )
    {
    }
    std::size_t begin() const { return _b; }
    std::size_t end() const { return _e; }
    const DeviceExecutionSpace& space() const { return _space; }

  private:
    DeviceExecutionSpace _space;
    std::size_t _b, _e;
};

namespace Impl
{
//! Kokkos::Profiling::ScopedRegion stand-in: an NVTX range with the reference's region name.
struct ScopedRegion
{
#ifdef CABANA_B200_HAVE_NVTX
    explicit ScopedRegion( const char* name ) { nvtxRangePushA( name ); }
    ~ScopedRegion() { nvtxRangePop(); }
#else
    explicit ScopedRegion( const char* ) {}
#endif
    ScopedRegion( const ScopedRegion& ) = delete;
    ScopedRegion& operator=( const ScopedRegion& ) = delete;
};
inline void check( int rc, const char* where )
{
    if ( rc != CB_OK )
        throw std::runtime_error( std::string( where ) + ": " + cb_last_error_string() );
}

} // namespace Impl
inline void DeviceExecutionSpace::fence() const
{
    Impl::check( cb_stream_synchronize( _stream ), "Cabana::DeviceExecutionSpace::fence" );
}
namespace Impl
{
//! Ref-counted device allocation (Kokkos::View ownership semantics: copies are shallow).
template <class T>
std::shared_ptr<T> device_alloc( std::size_t n )
{
    void* p = nullptr;
    check( cb_malloc( &p, sizeof( T ) * ( n ? n : 1 ) ), "cb_malloc" );
    return std::shared_ptr<T>( static_cast<T*>( p ), []( T* q ) { cb_free( q ); } );
}
} // namespace Impl

//---------------------------------------------------------------------------//
// Slice: unmanaged strided view of one AoSoA member
// (core/src/Cabana_Slice.hpp:134-140; element (i,d) = data[Stride*(i/V) + i%V + V*d]).
//---------------------------------------------------------------------------//
template <class T, int NumComp = 3>
class Slice
{
  public:
    using value_type = T;
    using memory_space = DeviceSpace;
    static constexpr int num_comp = NumComp;

    Slice() = default;
    Slice( T* data, std::size_t n, std::size_t stride, int vector_length )
        : _data( data )
        , _n( n )
        , _stride( stride )
        , _vlen( vector_length )
    {
    }
    std::size_t size() const { return _n; }
    T* data() const { return _data; }
    std::size_t stride( int ) const { return _stride; }
    int vectorLength() const { return _vlen; }

    CABANA_B200_FUNCTION T& operator()( std::size_t i, int d ) const
    {
        return _data[_stride * ( i / _vlen ) + ( i % _vlen ) + (std::size_t)_vlen * d];
    }

    cb_positions positions() const
    {
        static_assert( std::is_same<typename std::remove_const<T>::type, double>::value,
                       "positions are double (float positions are not instantiated)" );
        return cb_positions{ _data, (int64_t)_n, (int64_t)_stride, (int32_t)_vlen,
                             (int64_t)_vlen };
    }
    cb_field field() const
    {
        return cb_field{ (void*)_data,       (int64_t)_n,       (int64_t)_stride,
                         (int32_t)_vlen,     (int64_t)_vlen,    (int32_t)NumComp,
                         (int32_t)sizeof( T ) };
    }

  private:
    T* _data = nullptr;
    std::size_t _n = 0;
    std::size_t _stride = 0;
    int _vlen = 1;
};

//! Rank-2 (n, NumComp) device view with arbitrary strides (Kokkos::View<T**> stand-in).
template <class T, int NumComp = 3>
class View2D
{
  public:
    using value_type = T;
    using memory_space = DeviceSpace;
    static constexpr int num_comp = NumComp;
    View2D() = default;
    View2D( T* data, std::size_t n, std::size_t stride0 = NumComp, std::size_t stride1 = 1 )
        : _data( data )
        , _n( n )
        , _s0( stride0 )
        , _s1( stride1 )
    {
    }
    std::size_t size() const { return _n; }
    std::size_t extent( int d ) const { return d == 0 ? _n : NumComp; }
    T* data() const { return _data; }
    CABANA_B200_FUNCTION T& operator()( std::size_t i, int d ) const
    {
        return _data[_s0 * i + _s1 * d];
    }
    cb_positions positions() const
    {
        return cb_positions{ _data, (int64_t)_n, (int64_t)_s0, 1, (int64_t)_s1 };
    }
    cb_field field() const
    {
        return cb_field{ (void*)_data, (int64_t)_n,       (int64_t)_s0,        1,
                         (int64_t)_s1, (int32_t)NumComp, (int32_t)sizeof( T ) };
    }

  private:
    T* _data = nullptr;
    std::size_t _n = 0, _s0 = NumComp, _s1 = 1;
};

template <class P>
std::size_t size( const P& p )
{
    return p.size();
}

//---------------------------------------------------------------------------//
// Tags (Cabana_NeighborList.hpp:35-58, Cabana_VerletList.hpp:30-41,
// Cabana_Parallel.hpp:190-212)
//---------------------------------------------------------------------------//
class FullNeighborTag
{
};
class HalfNeighborTag
{
};
struct VerletLayoutCSR
{
};
struct VerletLayout2D
{
};
class FirstNeighborsTag
{
};
//! Second-neighbour (triplet) traversal tag (Cabana_Parallel.hpp:200-204)
class SecondNeighborsTag
{
};
class SerialOpTag
{
};
class TeamOpTag
{
};
class TeamVectorOpTag
{
};

namespace Impl
{
template <class Tag>
struct algorithm_enum;
template <>
struct algorithm_enum<FullNeighborTag>
{
    static constexpr int value = CB_NEIGHBOR_FULL;
};
template <>
struct algorithm_enum<HalfNeighborTag>
{
    static constexpr int value = CB_NEIGHBOR_HALF;
};
template <class Tag>
struct layout_enum;
template <>
struct layout_enum<VerletLayoutCSR>
{
    static constexpr int value = CB_LAYOUT_CSR;
};
template <>
struct layout_enum<VerletLayout2D>
{
    static constexpr int value = CB_LAYOUT_2D;
};
template <class Tag>
struct op_enum;
template <>
struct op_enum<SerialOpTag>
{
    static constexpr int value = CB_OP_SERIAL;
};
template <>
struct op_enum<TeamOpTag>
{
    static constexpr int value = CB_OP_TEAM;
};
template <>
struct op_enum<TeamVectorOpTag>
{
    static constexpr int value = CB_OP_TEAM_VECTOR;
};

template <class ArrayType>
std::array<double, 3> to_array3( const ArrayType& a )
{
    return { (double)a[0], (double)a[1], (double)a[2] };
}
} // namespace Impl

//---------------------------------------------------------------------------//
// BinningData (core/src/Cabana_Sort.hpp:37-136): a value type over the binning arrays.  The
// accessors read DEVICE memory, so (as in the reference, where they are KOKKOS_INLINE_FUNCTIONs
// over device Views) they are for device code, or for a host mirror of the arrays.
//---------------------------------------------------------------------------//
template <class MemorySpace>
class BinningData
{
  public:
    using memory_space = MemorySpace;
    using size_type = typename MemorySpace::size_type;

    BinningData() = default;
    BinningData( const std::size_t begin, const std::size_t end, const int* counts,
                 const size_type* offsets, const size_type* permute_vector, const int nbin )
        : _begin( begin )
        , _end( end )
        , _nbin( nbin )
        , _counts( counts )
        , _offsets( offsets )
        , _permute_vector( permute_vector )
    {
    }
    CABANA_B200_FUNCTION int numBin() const { return _nbin; }
    CABANA_B200_FUNCTION int binSize( const size_type bin_id ) const { return _counts[bin_id]; }
    CABANA_B200_FUNCTION size_type binOffset( const size_type bin_id ) const
    {
        return _offsets[bin_id];
    }
    CABANA_B200_FUNCTION size_type permutation( const size_type tuple_id ) const
    {
        return _permute_vector[tuple_id];
    }
    CABANA_B200_FUNCTION std::size_t rangeBegin() const { return _begin; }
    CABANA_B200_FUNCTION std::size_t rangeEnd() const { return _end; }
    const size_type* permuteData() const { return _permute_vector; }

  private:
    std::size_t _begin = 0, _end = 0;
    int _nbin = 0;
    const int* _counts = nullptr;
    const size_type* _offsets = nullptr;
    const size_type* _permute_vector = nullptr;
};

//! Device-side view of a LinkedCellList: the accessors of Cabana_LinkedCellList.hpp:489-650,
//! :841-872 that functors call (binSize/binOffset by ijk or cardinal, permutation,
//! getParticleBin, getParticle, getStencilCells, cardinalBinIndex/ijkBinIndex).
struct LinkedCellListView
{
    cb_lcl_view v;
    CABANA_B200_FUNCTION int totalBins() const { return (int)v.num_cells; }
    CABANA_B200_FUNCTION int numBin( const int dim ) const { return v.grid.nx[dim]; }
    CABANA_B200_FUNCTION int cardinalBinIndex( const int i, const int j, const int k ) const
    {
        return ( i * v.grid.nx[1] + j ) * v.grid.nx[2] + k;
    }
    CABANA_B200_FUNCTION void ijkBinIndex( const int cardinal, int& i, int& j, int& k ) const
    {
        i = cardinal / ( v.grid.nx[1] * v.grid.nx[2] );
        j = ( cardinal / v.grid.nx[2] ) % v.grid.nx[1];
        k = cardinal % v.grid.nx[2];
    }
    CABANA_B200_FUNCTION int binSize( const int i, const int j, const int k ) const
    {
        return v.counts[cardinalBinIndex( i, j, k )];
    }
    CABANA_B200_FUNCTION unsigned binOffset( const int i, const int j, const int k ) const
    {
        return v.offsets[cardinalBinIndex( i, j, k )];
    }
    CABANA_B200_FUNCTION unsigned permutation( const int particle_id ) const
    {
        return v.permute[particle_id];
    }
    CABANA_B200_FUNCTION std::size_t rangeBegin() const { return (std::size_t)v.begin; }
    CABANA_B200_FUNCTION std::size_t rangeEnd() const { return (std::size_t)v.end; }
    CABANA_B200_FUNCTION bool sorted() const { return v.sorted != 0; }
    CABANA_B200_FUNCTION int getParticleBin( const int particle_index ) const
    {
        return v.particle_bins[particle_index - v.begin];
    }
    CABANA_B200_FUNCTION std::size_t getParticle( const int offset ) const
    {
        return v.sorted ? (std::size_t)( offset + v.begin ) : (std::size_t)v.permute[offset];
    }
    CABANA_B200_FUNCTION void getStencilCells( const int cell, int& imin, int& imax, int& jmin,
                                               int& jmax, int& kmin, int& kmax ) const
    {
        const int ny = v.stencil_grid.nx[1], nz = v.stencil_grid.nx[2], R = v.cell_range;
        const int ci = cell / ( ny * nz ), cj = ( cell / nz ) % ny, ck = cell % nz;
        imin = ci - R > 0 ? ci - R : 0;
        imax = ci + R + 1 < v.stencil_grid.nx[0] ? ci + R + 1 : v.stencil_grid.nx[0];
        jmin = cj - R > 0 ? cj - R : 0;
        jmax = cj + R + 1 < ny ? cj + R + 1 : ny;
        kmin = ck - R > 0 ? ck - R : 0;
        kmax = ck + R + 1 < nz ? ck + R + 1 : nz;
    }
};

//---------------------------------------------------------------------------//
// LinkedCellList (core/src/Cabana_LinkedCellList.hpp:128-909)
//---------------------------------------------------------------------------//
template <class MemorySpace, class Scalar = double, std::size_t NumSpaceDim = 3>
class LinkedCellList
{
    static_assert( NumSpaceDim == 3 || NumSpaceDim == 2, "VerletList: 2-D or 3-D space" );
    static_assert( std::is_same<Scalar, double>::value, "float positions: SURVEY.md 8f (next)" );

  public:
    using memory_space = MemorySpace;
    using size_type = typename MemorySpace::size_type;
    static constexpr std::size_t num_space_dim = NumSpaceDim;

    LinkedCellList() = default;

    //! Simple constructor (:166-192)
    template <class PositionType, class ArrayType>
    LinkedCellList( PositionType positions, const ArrayType& grid_delta,
                    const ArrayType& grid_min, const ArrayType& grid_max )
    {
        create( grid_delta, grid_min, grid_max, -1.0, 1.0 );
        build( positions, 0, positions.size() );
    }
    //! Partial range constructor (:210-235)
    template <class PositionType, class ArrayType>
    LinkedCellList( PositionType positions, const std::size_t begin, const std::size_t end,
                    const ArrayType& grid_delta, const ArrayType& grid_min,
                    const ArrayType& grid_max )
    {
        create( grid_delta, grid_min, grid_max, -1.0, 1.0 );
        build( positions, begin, end );
    }
    //! Explicit stencil constructor (:249-277)
    template <class PositionType, class ArrayType>
    LinkedCellList( PositionType positions, const ArrayType& grid_delta,
                    const ArrayType& grid_min, const ArrayType& grid_max,
                    const Scalar neighborhood_radius, const Scalar cell_size_ratio = 1 )
    {
        create( grid_delta, grid_min, grid_max, neighborhood_radius, cell_size_ratio );
        build( positions, 0, positions.size() );
    }
    //! Explicit stencil and partial range constructor (:296-325)
    template <class PositionType, class ArrayType>
    LinkedCellList( PositionType positions, const std::size_t begin, const std::size_t end,
                    const ArrayType& grid_delta, const ArrayType& grid_min,
                    const ArrayType& grid_max, const Scalar neighborhood_radius,
                    const Scalar cell_size_ratio = 1 )
    {
        create( grid_delta, grid_min, grid_max, neighborhood_radius, cell_size_ratio );
        build( positions, begin, end );
    }

    //! build (:651-773)
    template <class PositionType>
    void build( PositionType positions, const std::size_t begin, const std::size_t end )
    {
        cb_positions x = positions.positions();
        Impl::check( cb_lcl_build( _h.get(), &x, (int64_t)begin, (int64_t)end, nullptr ),
                     "Cabana::LinkedCellList::build" );
        refresh();
    }
    template <class PositionType>
    void build( PositionType positions )
    {
        build( positions, 0, positions.size() );
    }

    int numParticles() const { return (int)( _v.end - _v.begin ); }
    std::size_t getParticleBegin() const { return (std::size_t)_v.begin; }
    std::size_t getParticleEnd() const { return (std::size_t)_v.end; }
    std::size_t rangeBegin() const { return (std::size_t)_v.begin; }
    std::size_t rangeEnd() const { return (std::size_t)_v.end; }
    int totalBins() const { return (int)_v.num_cells; }
    int numBin( const int dim ) const { return _v.grid.nx[dim]; }
    size_type cardinalBinIndex( const int i, const int j, const int k ) const
    {
        return (size_type)cb_grid_cardinal_cell_index( &_v.grid, i, j, k );
    }
    void ijkBinIndex( const int cardinal, int& i, int& j, int& k ) const
    {
        int32_t ijk[3];
        cb_grid_ijk_bin_index( &_v.grid, cardinal, ijk );
        i = ijk[0];
        j = ijk[1];
        k = ijk[2];
    }
    bool sorted() const { return _v.sorted != 0; }
    void update( const bool sorted )
    {
        Impl::check( cb_lcl_update( _h.get(), sorted ? 1 : 0 ), "LinkedCellList::update" );
        refresh();
    }
    void getStencilCells( const int cell, int& imin, int& imax, int& jmin, int& jmax,
                          int& kmin, int& kmax ) const
    {
        int32_t mn[3], mx[3];
        cb_stencil_get_cells( &_v.stencil_grid, _v.cell_range, cell, mn, mx );
        imin = mn[0];
        imax = mx[0];
        jmin = mn[1];
        jmax = mx[1];
        kmin = mn[2];
        kmax = mx[2];
    }

    //! Device pointers (what the reference's KOKKOS_INLINE accessors read).
    const int* binSizeData() const { return _v.counts; }
    const size_type* binOffsetData() const { return _v.offsets; }
    const size_type* permutationData() const { return _v.permute; }
    const int* getParticleBins() const { return _v.particle_bins; }

    //! Host copies for tests / set-up code (binSize, binOffset, permutation, bins).
    struct HostMirror
    {
        std::vector<int> counts;
        std::vector<size_type> offsets;
        std::vector<size_type> permute;
        std::vector<int> particle_bins;
    };
    HostMirror hostMirror() const
    {
        HostMirror m;
        const std::size_t nc = (std::size_t)_v.num_cells, np = (std::size_t)numParticles();
        m.counts.resize( nc );
        m.offsets.resize( nc + 1 );
        m.permute.resize( np );
        m.particle_bins.resize( np );
        cb_memcpy_d2h( m.counts.data(), _v.counts, nc * sizeof( int ), nullptr );
        cb_memcpy_d2h( m.offsets.data(), _v.offsets, ( nc + 1 ) * sizeof( size_type ), nullptr );
        cb_memcpy_d2h( m.permute.data(), _v.permute, np * sizeof( size_type ), nullptr );
        cb_memcpy_d2h( m.particle_bins.data(), _v.particle_bins, np * sizeof( int ), nullptr );
        Impl::check( cb_stream_synchronize( nullptr ), "LinkedCellList::hostMirror" );
        return m;
    }

    cb_lcl* handle() const { return _h.get(); }
    void refresh() { Impl::check( cb_lcl_get( _h.get(), &_v ), "cb_lcl_get" ); }
    //! POD copy (device pointers + grids) that kernels and functors take by value
    using device_view_type = LinkedCellListView;
    LinkedCellListView deviceView() const { return LinkedCellListView{ _v }; }
    //! binningData() (:745-749): 1-D binning data over the cells
    BinningData<MemorySpace> binningData() const
    {
        return BinningData<MemorySpace>( (std::size_t)_v.begin, (std::size_t)_v.end, _v.counts,
                                         _v.offsets, _v.permute, (int)_v.num_cells );
    }

  private:
    template <class ArrayType>
    void create( const ArrayType& delta, const ArrayType& mn, const ArrayType& mx,
                 double radius, double ratio )
    {
        auto d = Impl::to_array3( delta ), a = Impl::to_array3( mn ), b = Impl::to_array3( mx );
        cb_lcl* raw = nullptr;
        Impl::check( cb_lcl_create( &raw, d.data(), a.data(), b.data(), radius, ratio ),
                     "Cabana::LinkedCellList" );
        _h = std::shared_ptr<cb_lcl>( raw, []( cb_lcl* p ) { cb_lcl_destroy( p ); } );
    }

    std::shared_ptr<cb_lcl> _h; // shallow copies share the list, like Kokkos::View members
    cb_lcl_view _v{};
};

//! createLinkedCellList overloads (:915-1095)
template <class PositionType, class ArrayType>
auto createLinkedCellList( PositionType positions, const ArrayType& grid_delta,
                           const ArrayType& grid_min, const ArrayType& grid_max )
{
    return LinkedCellList<typename PositionType::memory_space, double, 3>(
        positions, grid_delta, grid_min, grid_max );
}
template <class PositionType, class ArrayType>
auto createLinkedCellList( PositionType positions, const std::size_t begin,
                           const std::size_t end, const ArrayType& grid_delta,
                           const ArrayType& grid_min, const ArrayType& grid_max )
{
    return LinkedCellList<typename PositionType::memory_space, double, 3>(
        positions, begin, end, grid_delta, grid_min, grid_max );
}
template <class PositionType, class ArrayType>
auto createLinkedCellList( PositionType positions, const ArrayType& grid_delta,
                           const ArrayType& grid_min, const ArrayType& grid_max,
                           const double neighborhood_radius,
                           const double cell_size_ratio = 1.0 )
{
    return LinkedCellList<typename PositionType::memory_space, double, 3>(
        positions, grid_delta, grid_min, grid_max, neighborhood_radius, cell_size_ratio );
}

template <class T>
struct is_linked_cell_list : public std::false_type
{
};
template <class M, class S, std::size_t D>
struct is_linked_cell_list<LinkedCellList<M, S, D>> : public std::true_type
{
};

//! permute(BinningData, slice...) (core/src/Cabana_Sort.hpp:549-715): tmp[i] = data[perm[i]],
//! data[begin+i] = tmp[i] for every member passed.
template <class MemorySpace, class... Members>
void permute( const BinningData<MemorySpace>& binning_data, Members&... members )
{
    cb_field f[] = { members.field()... };
    Impl::check( cb_binning_permute( (int64_t)binning_data.rangeBegin(),
                                     (int64_t)binning_data.rangeEnd(),
                                     binning_data.permuteData(), f,
                                     (int)sizeof...( Members ), nullptr ),
                 "Cabana::permute" );
}

//! permute(LinkedCellList&, slice...) (:1130-1145).  Pass every member slice of the
//! AoSoA to permute the whole AoSoA.
template <class LinkedCellListType, class... Members>
typename std::enable_if<is_linked_cell_list<LinkedCellListType>::value>::type
permute( LinkedCellListType& linked_cell_list, Members&... members )
{
    cb_field f[] = { members.field()... };
    Impl::check( cb_lcl_permute( linked_cell_list.handle(), f, (int)sizeof...( Members ), nullptr ),
                 "Cabana::permute" );
    linked_cell_list.refresh();
}

//---------------------------------------------------------------------------//
// VerletListData (Cabana_VerletList.hpp:50-113).  counts/offsets/neighbors are device
// pointers owned by the list handle (shallow copies share them).
//---------------------------------------------------------------------------//
template <class MemorySpace, class LayoutTag>
struct VerletListData;

template <class MemorySpace>
struct VerletListData<MemorySpace, VerletLayoutCSR>
{
    using memory_space = MemorySpace;
    int* counts = nullptr;
    int* offsets = nullptr;
    int* neighbors = nullptr;
    std::size_t num_particles = 0;
    std::size_t total = 0; // neighbors.extent(0)
    std::size_t max_n = 0;
    CABANA_B200_FUNCTION std::size_t num( std::size_t i ) const { return counts[i]; }
    CABANA_B200_FUNCTION std::size_t get( std::size_t i, std::size_t n ) const
    {
        return neighbors[offsets[i] + n];
    }
    CABANA_B200_FUNCTION void setNeighbor( int pid, int nid, int new_id ) const
    {
        neighbors[offsets[pid] + nid] = new_id;
    }
};

template <class MemorySpace>
struct VerletListData<MemorySpace, VerletLayout2D>
{
    using memory_space = MemorySpace;
    int* counts = nullptr;
    int* neighbors = nullptr; // (i,n) at neighbors[i*row_stride + n]
    std::size_t num_particles = 0;
    std::size_t row_stride = 0; // neighbors.extent(1)
    std::size_t total = 0;
    std::size_t max_n = 0;
    CABANA_B200_FUNCTION std::size_t num( std::size_t i ) const { return counts[i]; }
    CABANA_B200_FUNCTION std::size_t get( std::size_t i, std::size_t n ) const
    {
        return neighbors[i * row_stride + n];
    }
    CABANA_B200_FUNCTION void setNeighbor( int pid, int nid, int new_id ) const
    {
        neighbors[pid * row_stride + nid] = new_id;
    }
};

//! Trivially-copyable device-side view of a VerletList: what kernels capture by value
//! (the reference captures the list itself; its Views are device handles already).
template <class MemorySpace, class AlgorithmTag, class LayoutTag>
struct VerletListView
{
    using memory_space = MemorySpace;
    using algorithm_tag = AlgorithmTag;
    using layout_tag = LayoutTag;
    VerletListData<MemorySpace, LayoutTag> _data;
    CABANA_B200_FUNCTION void setNeighbor( std::size_t i, std::size_t n, int new_index ) const
    {
        _data.setNeighbor( (int)i, (int)n, new_index );
    }
};

//---------------------------------------------------------------------------//
// VerletList (Cabana_VerletList.hpp:824-1495)
//---------------------------------------------------------------------------//
template <class MemorySpace, class AlgorithmTag, class LayoutTag,
          class BuildTag = TeamVectorOpTag, std::size_t NumSpaceDim = 3>
class VerletList
{
    static_assert( NumSpaceDim == 3 || NumSpaceDim == 2, "VerletList: 2-D or 3-D space" );

  public:
    static constexpr std::size_t num_space_dim = NumSpaceDim;
    using memory_space = MemorySpace;
    using algorithm_tag = AlgorithmTag;
    using layout_tag = LayoutTag;

    //! Verlet list data (public in the reference too, :841).
    VerletListData<memory_space, LayoutTag> _data;

    VerletList() = default;

    //! Constructors (:875-1017, with a leading ExecutionSpace :1260-1281) and build (:1351-1392).
    template <class PositionType, class ArrayType>
    VerletList( PositionType x, const std::size_t begin, const std::size_t end,
                const double neighborhood_radius, const double cell_size_ratio,
                const ArrayType& grid_min, const ArrayType& grid_max,
                const std::size_t max_neigh = 0 )
    {
        build( DeviceExecutionSpace(), x, begin, end, neighborhood_radius, cell_size_ratio,
               grid_min, grid_max, max_neigh );
    }
    template <class PositionType, class ArrayType>
    VerletList( const DeviceExecutionSpace& exec_space, PositionType x, const std::size_t begin,
                const std::size_t end, const double neighborhood_radius,
                const double cell_size_ratio, const ArrayType& grid_min,
                const ArrayType& grid_max, const std::size_t max_neigh = 0 )
    {
        build( exec_space, x, begin, end, neighborhood_radius, cell_size_ratio, grid_min,
               grid_max, max_neigh );
    }
    //! Per-particle cutoff: background radius for the linked cells + a radius slice/view (:989-1017)
    template <class PositionType, class RadiusType, class ArrayType>
    VerletList( PositionType x, const std::size_t begin, const std::size_t end,
                const double background_radius, RadiusType neighborhood_radius,
                const double cell_size_ratio, const ArrayType& grid_min,
                const ArrayType& grid_max, const std::size_t max_neigh = 0,
                decltype( std::declval<RadiusType>().field() )* = nullptr )
    {
        build( DeviceExecutionSpace(), x, begin, end, background_radius, neighborhood_radius,
               cell_size_ratio, grid_min, grid_max, max_neigh );
    }
    template <class PositionType, class RadiusType, class ArrayType>
    VerletList( const DeviceExecutionSpace& exec_space, PositionType x, const std::size_t begin,
                const std::size_t end, const double background_radius,
                RadiusType neighborhood_radius, const double cell_size_ratio,
                const ArrayType& grid_min, const ArrayType& grid_max,
                const std::size_t max_neigh = 0,
                decltype( std::declval<RadiusType>().field() )* = nullptr )
    {
        build( exec_space, x, begin, end, background_radius, neighborhood_radius,
               cell_size_ratio, grid_min, grid_max, max_neigh );
    }

    template <class PositionType, class ArrayType>
    void build( PositionType x, const std::size_t begin, const std::size_t end,
                const double neighborhood_radius, const double cell_size_ratio,
                const ArrayType& grid_min, const ArrayType& grid_max,
                const std::size_t max_neigh = 0 )
    {
        build( DeviceExecutionSpace(), x, begin, end, neighborhood_radius, cell_size_ratio,
               grid_min, grid_max, max_neigh );
    }
    template <class PositionType, class ArrayType>
    void build( const DeviceExecutionSpace& exec_space, PositionType x, const std::size_t begin,
                const std::size_t end, const double neighborhood_radius,
                const double cell_size_ratio, const ArrayType& grid_min,
                const ArrayType& grid_max, const std::size_t max_neigh = 0 )
    {
        ensure_handle();
        cb_positions xd = x.positions();
        if constexpr ( NumSpaceDim == 2 )
        {
            // NumSpaceDim = 2 (:377-392, :626-639): two-component positions and grid bounds
            const double mn[2] = { (double)grid_min[0], (double)grid_min[1] };
            const double mx[2] = { (double)grid_max[0], (double)grid_max[1] };
            Impl::check( cb_verlet_build_2d( _h.get(), &xd, (int64_t)begin, (int64_t)end,
                                             neighborhood_radius, cell_size_ratio, mn, mx,
                                             (int64_t)max_neigh,
                                             Impl::algorithm_enum<AlgorithmTag>::value,
                                             Impl::layout_enum<LayoutTag>::value,
                                             Impl::op_enum<BuildTag>::value,
                                             exec_space.stream() ),
                         "Cabana::VerletList::build" );
        }
        else
        {
            auto mn = Impl::to_array3( grid_min ), mx = Impl::to_array3( grid_max );
            Impl::check( cb_verlet_build( _h.get(), &xd, (int64_t)begin, (int64_t)end,
                                          neighborhood_radius, cell_size_ratio, mn.data(),
                                          mx.data(), (int64_t)max_neigh,
                                          Impl::algorithm_enum<AlgorithmTag>::value,
                                          Impl::layout_enum<LayoutTag>::value,
                                          Impl::op_enum<BuildTag>::value,
                                          exec_space.stream() ),
                         "Cabana::VerletList::build" );
        }
        Impl::check( cb_verlet_get( _h.get(), &_view ), "cb_verlet_get" );
        fill_data( _data );
    }
    template <class PositionType, class RadiusType, class ArrayType>
    auto build( PositionType x, const std::size_t begin, const std::size_t end,
                const double background_radius, RadiusType neighborhood_radius,
                const double cell_size_ratio, const ArrayType& grid_min,
                const ArrayType& grid_max, const std::size_t max_neigh = 0 )
        -> decltype( neighborhood_radius.field(), void() )
    {
        build( DeviceExecutionSpace(), x, begin, end, background_radius, neighborhood_radius,
               cell_size_ratio, grid_min, grid_max, max_neigh );
    }
    template <class PositionType, class RadiusType, class ArrayType>
    auto build( const DeviceExecutionSpace& exec_space, PositionType x, const std::size_t begin,
                const std::size_t end, const double background_radius,
                RadiusType neighborhood_radius, const double cell_size_ratio,
                const ArrayType& grid_min, const ArrayType& grid_max,
                const std::size_t max_neigh = 0 )
        -> decltype( neighborhood_radius.field(), void() )
    {
        ensure_handle();
        auto mn = Impl::to_array3( grid_min ), mx = Impl::to_array3( grid_max );
        cb_positions xd = x.positions();
        cb_field rd = neighborhood_radius.field();
        Impl::check( cb_verlet_build_radii( _h.get(), &xd, &rd, (int64_t)begin, (int64_t)end,
                                            background_radius, cell_size_ratio, mn.data(),
                                            mx.data(), (int64_t)max_neigh,
                                            Impl::algorithm_enum<AlgorithmTag>::value,
                                            Impl::layout_enum<LayoutTag>::value,
                                            Impl::op_enum<BuildTag>::value, exec_space.stream() ),
                     "Cabana::VerletList::build" );
        Impl::check( cb_verlet_get( _h.get(), &_view ), "cb_verlet_get" );
        fill_data( _data );
    }

    //! setNeighbor (:1488-1494) from the host; inside kernels use _data.setNeighbor.
    void setNeighbor( const std::size_t particle_index, const std::size_t neighbor_index,
                      const int new_index ) const
    {
        Impl::check( cb_verlet_set_neighbor( _h.get(), (int64_t)particle_index,
                                             (int64_t)neighbor_index, new_index, nullptr ),
                     "Cabana::VerletList::setNeighbor" );
    }

    const cb_verlet_view& view() const { return _view; }

    //! Hook for fused steps that build this list through another C entry of the library
    //! (Cabana::SlabPeerHalo::step -> cb_slab_step): fn( cb_verlet*, algorithm, layout, build_op )
    //! does the build, then the public data is refreshed as after build().
    template <class BuildFunction>
    void buildThrough( BuildFunction&& fn )
    {
        static_assert( NumSpaceDim == 3, "fused builds are three-dimensional" );
        ensure_handle();
        fn( _h.get(), Impl::algorithm_enum<AlgorithmTag>::value, Impl::layout_enum<LayoutTag>::value,
            Impl::op_enum<BuildTag>::value );
        Impl::check( cb_verlet_get( _h.get(), &_view ), "cb_verlet_get" );
        fill_data( _data );
    }

    using device_view_type = VerletListView<MemorySpace, AlgorithmTag, LayoutTag>;
    device_view_type deviceView() const { return device_view_type{ _data }; }

  private:
    void ensure_handle()
    {
        if ( !_h )
        {
            cb_verlet* raw = nullptr;
            Impl::check( cb_verlet_create( &raw ), "Cabana::VerletList" );
            _h = std::shared_ptr<cb_verlet>( raw, []( cb_verlet* p ) { cb_verlet_destroy( p ); } );
        }
    }
    void fill_data( VerletListData<memory_space, VerletLayoutCSR>& d )
    {
        d.counts = _view.counts;
        d.offsets = _view.offsets;
        d.neighbors = _view.neighbors;
        d.num_particles = (std::size_t)_view.n;
        d.total = (std::size_t)_view.total;
        d.max_n = (std::size_t)_view.max_n;
    }
    void fill_data( VerletListData<memory_space, VerletLayout2D>& d )
    {
        d.counts = _view.counts;
        d.neighbors = _view.neighbors;
        d.num_particles = (std::size_t)_view.n;
        d.row_stride = (std::size_t)_view.row_stride;
        d.total = (std::size_t)_view.total;
        d.max_n = (std::size_t)_view.max_n;
    }

    std::shared_ptr<cb_verlet> _h;
    cb_verlet_view _view{};
};

//! createVerletList (:1528-1597), fixed and per-particle radius, with/without ExecutionSpace
template <class AlgorithmTag, class LayoutTag, class BuildTag, class PositionType,
          class ArrayType>
auto createVerletList( PositionType positions, const std::size_t begin,
                       const std::size_t end, const double radius,
                       const double cell_size_ratio, const ArrayType& grid_min,
                       const ArrayType& grid_max, const std::size_t max_neigh = 0 )
{
    return VerletList<typename PositionType::memory_space, AlgorithmTag, LayoutTag, BuildTag, 3>(
        positions, begin, end, radius, cell_size_ratio, grid_min, grid_max, max_neigh );
}
template <class AlgorithmTag, class LayoutTag, class BuildTag, class PositionType,
          class ArrayType>
auto createVerletList( const DeviceExecutionSpace& exec_space, PositionType positions,
                       const std::size_t begin, const std::size_t end, const double radius,
                       const double cell_size_ratio, const ArrayType& grid_min,
                       const ArrayType& grid_max, const std::size_t max_neigh = 0 )
{
    return VerletList<typename PositionType::memory_space, AlgorithmTag, LayoutTag, BuildTag, 3>(
        exec_space, positions, begin, end, radius, cell_size_ratio, grid_min, grid_max,
        max_neigh );
}
template <class AlgorithmTag, class LayoutTag, class BuildTag, class PositionType,
          class RadiusType, class ArrayType>
auto createVerletList( PositionType positions, const std::size_t begin,
                       const std::size_t end, const double background_radius,
                       RadiusType radius, const double cell_size_ratio,
                       const ArrayType& grid_min, const ArrayType& grid_max,
                       const std::size_t max_neigh = 0,
                       decltype( std::declval<RadiusType>().field() )* = nullptr )
{
    return VerletList<typename PositionType::memory_space, AlgorithmTag, LayoutTag, BuildTag, 3>(
        positions, begin, end, background_radius, radius, cell_size_ratio, grid_min, grid_max,
        max_neigh );
}

//---------------------------------------------------------------------------//
// NeighborList traits (Cabana_NeighborList.hpp:221-252; VerletList specialisations
// Cabana_VerletList.hpp:1603-1698).  numNeighbor / getNeighbor are callable in device
// code on a list captured BY VALUE (the data struct holds raw device pointers);
// totalNeighbor / maxNeighbor are O(1) here (computed once by the build) where the
// reference loops over all particles.
//---------------------------------------------------------------------------//
template <class NeighborListType>
class NeighborList;

template <class MemorySpace, class AlgorithmTag, class LayoutTag, class BuildTag,
          std::size_t Dim>
class NeighborList<VerletList<MemorySpace, AlgorithmTag, LayoutTag, BuildTag, Dim>>
{
  public:
    using memory_space = MemorySpace;
    using list_type = VerletList<MemorySpace, AlgorithmTag, LayoutTag, BuildTag, Dim>;

    static std::size_t totalNeighbor( const list_type& list ) { return list._data.total; }
    static std::size_t maxNeighbor( const list_type& list ) { return list._data.max_n; }
    //! Host-side element access (one small device read each; for tests and set-up code).
    static std::size_t numNeighbor( const list_type& list, const std::size_t particle_index )
    {
        int c = 0;
        cb_memcpy_d2h( &c, list._data.counts + particle_index, sizeof( int ), nullptr );
        Impl::check( cb_stream_synchronize( nullptr ), "NeighborList::numNeighbor" );
        return (std::size_t)c;
    }
    static std::size_t getNeighbor( const list_type& list, const std::size_t particle_index,
                                    const std::size_t neighbor_index )
    {
        const cb_verlet_view& v = list.view();
        long long at = 0;
        if ( v.layout == CB_LAYOUT_CSR )
        {
            int off = 0;
            cb_memcpy_d2h( &off, v.offsets + particle_index, sizeof( int ), nullptr );
            Impl::check( cb_stream_synchronize( nullptr ), "NeighborList::getNeighbor" );
            at = (long long)off + (long long)neighbor_index;
        }
        else
            at = (long long)particle_index * v.row_stride + (long long)neighbor_index * v.col_stride;
        int id = 0;
        cb_memcpy_d2h( &id, v.neighbors + at, sizeof( int ), nullptr );
        Impl::check( cb_stream_synchronize( nullptr ), "NeighborList::getNeighbor" );
        return (std::size_t)id;
    }
};

//! Device-side traits over the trivially-copyable view (used inside kernels).
template <class MemorySpace, class AlgorithmTag, class LayoutTag>
class NeighborList<VerletListView<MemorySpace, AlgorithmTag, LayoutTag>>
{
  public:
    using memory_space = MemorySpace;
    using list_type = VerletListView<MemorySpace, AlgorithmTag, LayoutTag>;
    CABANA_B200_FUNCTION static std::size_t totalNeighbor( const list_type& list )
    {
        return list._data.total;
    }
    CABANA_B200_FUNCTION static std::size_t maxNeighbor( const list_type& list )
    {
        return list._data.max_n;
    }
    CABANA_B200_FUNCTION static std::size_t numNeighbor( const list_type& list,
                                                        const std::size_t particle_index )
    {
        return list._data.num( particle_index );
    }
    CABANA_B200_FUNCTION static std::size_t getNeighbor( const list_type& list,
                                                        const std::size_t particle_index,
                                                        const std::size_t neighbor_index )
    {
        return list._data.get( particle_index, neighbor_index );
    }
};

//! NeighborList traits of a LinkedCellList (Cabana_LinkedCellList.hpp:1149-1303): the
//! "neighbours" of a particle are ALL particles of the stencil cells of its bin (itself
//! included, no cutoff) in stencil order; device code takes the list as its device view.
template <>
class NeighborList<LinkedCellListView>
{
  public:
    using list_type = LinkedCellListView;
    static constexpr std::size_t num_space_dim = 3;

    CABANA_B200_FUNCTION static std::size_t numNeighbor( const list_type& list,
                                                         const std::size_t particle_index )
    {
        int imin, imax, jmin, jmax, kmin, kmax;
        list.getStencilCells( list.getParticleBin( (int)particle_index ), imin, imax, jmin,
                              jmax, kmin, kmax );
        int total_count = 0;
        for ( int i = imin; i < imax; ++i )
            for ( int j = jmin; j < jmax; ++j )
                for ( int k = kmin; k < kmax; ++k )
                    total_count += list.binSize( i, j, k );
        return (std::size_t)total_count;
    }
    CABANA_B200_FUNCTION static std::size_t getNeighbor( const list_type& list,
                                                         const std::size_t particle_index,
                                                         const std::size_t neighbor_index )
    {
        int imin, imax, jmin, jmax, kmin, kmax;
        list.getStencilCells( list.getParticleBin( (int)particle_index ), imin, imax, jmin,
                              jmax, kmin, kmax );
        std::size_t total_count = 0, previous_count = 0;
        for ( int i = imin; i < imax; ++i )
            for ( int j = jmin; j < jmax; ++j )
                for ( int k = kmin; k < kmax; ++k )
                {
                    total_count += (std::size_t)list.binSize( i, j, k );
                    if ( total_count > neighbor_index ) // this neighbour is in this bin
                        return list.getParticle( (int)( list.binOffset( i, j, k ) +
                                                        ( neighbor_index - previous_count ) ) );
                    previous_count = total_count;
                }
        return 0; // (never reached for neighbor_index < numNeighbor)
    }
    CABANA_B200_FUNCTION static std::size_t totalNeighbor( const list_type& list )
    {
        std::size_t total_n = 0;
        for ( std::size_t p = list.rangeBegin(); p < list.rangeEnd(); ++p )
            total_n += numNeighbor( list, p );
        return total_n;
    }
    CABANA_B200_FUNCTION static std::size_t maxNeighbor( const list_type& list )
    {
        std::size_t max_n = 0;
        for ( std::size_t p = list.rangeBegin(); p < list.rangeEnd(); ++p )
        {
            const std::size_t c = numNeighbor( list, p );
            max_n = c > max_n ? c : max_n;
        }
        return max_n;
    }
};

//---------------------------------------------------------------------------//
// neighborHistogram (core/src/Cabana_NeighborList.hpp:283-332): histogram of neighbours per
// particle.  Row b = { upper bin edge, particles in the bin }.  The reference bins the counts with
// Cabana::binByKey (Kokkos::BinOp1D: mul = nbin/(max-min), bin = int(mul*(key-min))); this is
// set-up / diagnostics code, so the counts are brought to the host and binned there.
//---------------------------------------------------------------------------//
template <class ListType>
std::vector<std::array<int, 2>> neighborHistogram( const std::size_t num_particles,
                                                   const ListType& list, const int num_bin )
{
    std::vector<int> num_neigh( num_particles );
    if ( num_particles > 0 )
    {
        cb_memcpy_d2h( num_neigh.data(), list._data.counts, num_particles * sizeof( int ),
                       nullptr );
        Impl::check( cb_stream_synchronize( nullptr ), "Cabana::neighborHistogram" );
    }
    int kmin = 0, kmax = 0;
    if ( num_particles > 0 )
    {
        kmin = *std::min_element( num_neigh.begin(), num_neigh.end() );
        kmax = *std::max_element( num_neigh.begin(), num_neigh.end() );
    }
    std::vector<int> bin_size( num_bin + 1, 0 );
    const double mul = kmax > kmin ? (double)num_bin / (double)( kmax - kmin ) : 0.0;
    for ( int c : num_neigh )
        ++bin_size[(int)( mul * (double)( c - kmin ) )];
    const int max_neigh = (int)NeighborList<ListType>::maxNeighbor( list );
    double bin_width = (double)max_neigh / (double)num_bin;
    if ( num_bin > max_neigh )
        bin_width = 1;
    std::vector<std::array<int, 2>> histogram( num_bin );
    for ( int b = 0; b < num_bin; ++b )
        histogram[b] = { (int)( ( b + 1 ) * bin_width ), bin_size[b] };
    return histogram;
}

//---------------------------------------------------------------------------//
// Pre-compiled consumers (C ABI): the Lennard-Jones functor of the benchmark.
//---------------------------------------------------------------------------//
namespace B200
{
//! neighbor_parallel_for with the LJ functor; OpTag = SerialOpTag or TeamOpTag.
template <class ListType, class PositionType, class ForceType, class OpTag>
void neighbor_parallel_for_lj( const RangePolicy<>& policy, const ListType& list,
                               PositionType x, ForceType f, double eps, double sigma,
                               double rc, OpTag )
{
    const bool newton =
        std::is_same<typename ListType::algorithm_tag, HalfNeighborTag>::value;
    cb_positions xd = x.positions();
    cb_field fd = f.field();
    Impl::check( cb_neighbor_for_lj( &list.view(), &xd, &fd, eps, sigma, rc, newton ? 1 : 0,
                                     Impl::op_enum<OpTag>::value, (int64_t)policy.begin(),
                                     (int64_t)policy.end(), policy.space().stream() ),
                 "Cabana::neighbor_parallel_for(LJ)" );
}
//! neighbor_parallel_reduce with the LJ pair energy.
template <class ListType, class PositionType, class OpTag>
double neighbor_parallel_reduce_lj( const RangePolicy<>& policy, const ListType& list,
                                    PositionType x, double eps, double sigma, double rc,
                                    OpTag )
{
    const double scale =
        std::is_same<typename ListType::algorithm_tag, HalfNeighborTag>::value ? 1.0 : 0.5;
    cb_positions xd = x.positions();
    double e = 0.0;
    Impl::check( cb_neighbor_reduce_lj( &list.view(), &xd, eps, sigma, rc, scale,
                                        Impl::op_enum<OpTag>::value, (int64_t)policy.begin(),
                                        (int64_t)policy.end(), &e, policy.space().stream() ),
                 "Cabana::neighbor_parallel_reduce(LJ)" );
    return e;
}
} // namespace B200

#if defined( __CUDACC__ )
//---------------------------------------------------------------------------//
// Generic functors (nvcc only): the same loop shapes as Cabana_Parallel.hpp.
//---------------------------------------------------------------------------//
namespace Impl
{
// functorTagDispatch (Cabana_Parallel.hpp:39-78)
template <class WorkTag, class FunctorType, class... Args>
CABANA_B200_DEVICE typename std::enable_if<std::is_same<WorkTag, void>::value>::type
functorTagDispatch( const FunctorType& functor, Args&&... args )
{
    functor( static_cast<Args&&>( args )... );
}
template <class WorkTag, class FunctorType, class... Args>
CABANA_B200_DEVICE typename std::enable_if<!std::is_same<WorkTag, void>::value>::type
functorTagDispatch( const FunctorType& functor, Args&&... args )
{
    const WorkTag t{};
    functor( t, static_cast<Args&&>( args )... );
}

// Serial: one thread per particle, thread-local loop over neighbours (:280-288)
template <class WorkTag, class FunctorType, class ListType>
__global__ void k_neighbor_for_serial( FunctorType functor, ListType list, int begin, int end )
{
    using traits = NeighborList<ListType>;
    for ( int i = begin + blockIdx.x * blockDim.x + threadIdx.x; i < end;
          i += gridDim.x * blockDim.x )
    {
        const int nn = (int)traits::numNeighbor( list, i );
        for ( int n = 0; n < nn; ++n )
            functorTagDispatch<WorkTag>( functor, i, (int)traits::getNeighbor( list, i, n ) );
    }
}
// Team: one warp per particle, lanes stride the neighbours (:416-430)
template <class WorkTag, class FunctorType, class ListType>
__global__ void k_neighbor_for_team( FunctorType functor, ListType list, int begin, int end )
{
    using traits = NeighborList<ListType>;
    const int lane = threadIdx.x & 31;
    const int warp = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
    const int nwarps = ( gridDim.x * blockDim.x ) >> 5;
    for ( int i = begin + warp; i < end; i += nwarps )
    {
        const int nn = (int)traits::numNeighbor( list, i );
        for ( int n = lane; n < nn; n += 32 )
            functorTagDispatch<WorkTag>( functor, i, (int)traits::getNeighbor( list, i, n ) );
    }
}
// Sum of the threads' partial results: tree reduction of the CTA in shared memory, one value
// per CTA written to partials[blockIdx.x] (any ReduceType with += ; no atomics, deterministic).
template <class ReduceType>
CABANA_B200_DEVICE void block_reduce_store( const ReduceType& local, ReduceType* partials )
{
    extern __shared__ __align__( 16 ) unsigned char cb_reduce_smem[];
    ReduceType* sm = reinterpret_cast<ReduceType*>( cb_reduce_smem );
    sm[threadIdx.x] = local;
    __syncthreads();
    for ( int s = blockDim.x >> 1; s > 0; s >>= 1 )
    {
        if ( (int)threadIdx.x < s )
            sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if ( threadIdx.x == 0 )
        partials[blockIdx.x] = sm[0];
}
template <class ReduceType>
__global__ void k_sum_partials( const ReduceType* partials, int n, ReduceType* result )
{
    extern __shared__ __align__( 16 ) unsigned char cb_reduce_smem[];
    ReduceType* sm = reinterpret_cast<ReduceType*>( cb_reduce_smem );
    ReduceType local = ReduceType();
    for ( int i = threadIdx.x; i < n; i += blockDim.x )
        local += partials[i];
    sm[threadIdx.x] = local;
    __syncthreads();
    for ( int s = blockDim.x >> 1; s > 0; s >>= 1 )
    {
        if ( (int)threadIdx.x < s )
            sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if ( threadIdx.x == 0 )
        *result = sm[0];
}
template <class WorkTag, class FunctorType, class ListType, class ReduceType, bool Team>
__global__ void k_neighbor_reduce( FunctorType functor, ListType list, int begin, int end,
                                   ReduceType* partials )
{
    using traits = NeighborList<ListType>;
    ReduceType local = ReduceType();
    if ( Team )
    {
        const int lane = threadIdx.x & 31;
        const int warp = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
        const int nwarps = ( gridDim.x * blockDim.x ) >> 5;
        for ( int i = begin + warp; i < end; i += nwarps )
        {
            const int nn = (int)traits::numNeighbor( list, i );
            for ( int n = lane; n < nn; n += 32 )
                functorTagDispatch<WorkTag>( functor, i,
                                             (int)traits::getNeighbor( list, i, n ), local );
        }
    }
    else
    {
        for ( int i = begin + blockIdx.x * blockDim.x + threadIdx.x; i < end;
              i += gridDim.x * blockDim.x )
        {
            const int nn = (int)traits::numNeighbor( list, i );
            for ( int n = 0; n < nn; ++n )
                functorTagDispatch<WorkTag>( functor, i,
                                             (int)traits::getNeighbor( list, i, n ), local );
        }
    }
    block_reduce_store( local, partials );
}
// SecondNeighborsTag: every unordered pair (j,k) of neighbours of i, j before k in the row
// (Cabana_Parallel.hpp:315-365 Serial, :458-522 Team, :527-594 TeamVector)
template <class WorkTag, class FunctorType, class ListType>
__global__ void k_second_neighbor_for_serial( FunctorType functor, ListType list, int begin,
                                              int end )
{
    using traits = NeighborList<ListType>;
    for ( int i = begin + blockIdx.x * blockDim.x + threadIdx.x; i < end;
          i += gridDim.x * blockDim.x )
    {
        const int nn = (int)traits::numNeighbor( list, i );
        for ( int n = 0; n < nn; ++n )
        {
            const int j = (int)traits::getNeighbor( list, i, n );
            for ( int a = n + 1; a < nn; ++a )
                functorTagDispatch<WorkTag>( functor, i, j,
                                             (int)traits::getNeighbor( list, i, a ) );
        }
    }
}
// Team / TeamVector: one warp per particle; VECTOR = false: lanes stride the first neighbour
// and walk the second serially (:497-515); VECTOR = true: the warp strides the flattened
// (first, second) pairs so short rows still fill the lanes (:566-587)
template <class WorkTag, class FunctorType, class ListType, bool VECTOR>
__global__ void k_second_neighbor_for_team( FunctorType functor, ListType list, int begin,
                                            int end )
{
    using traits = NeighborList<ListType>;
    const int lane = threadIdx.x & 31;
    const int warp = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
    const int nwarps = ( gridDim.x * blockDim.x ) >> 5;
    for ( int i = begin + warp; i < end; i += nwarps )
    {
        const int nn = (int)traits::numNeighbor( list, i );
        if ( VECTOR )
        {
            // pair index p -> (n, a), n < a: row n of the strict upper triangle holds nn-1-n pairs
            const long long npairs = (long long)nn * ( nn - 1 ) / 2;
            int n = 0;
            long long row0 = 0; // first pair index of row n
            for ( long long p = lane; p < npairs; p += 32 )
            {
                while ( p >= row0 + ( nn - 1 - n ) )
                {
                    row0 += nn - 1 - n;
                    ++n;
                }
                const int a = n + 1 + (int)( p - row0 );
                functorTagDispatch<WorkTag>( functor, i, (int)traits::getNeighbor( list, i, n ),
                                             (int)traits::getNeighbor( list, i, a ) );
            }
        }
        else
        {
            for ( int n = lane; n < nn; n += 32 )
            {
                const int j = (int)traits::getNeighbor( list, i, n );
                for ( int a = n + 1; a < nn; ++a )
                    functorTagDispatch<WorkTag>( functor, i, j,
                                                 (int)traits::getNeighbor( list, i, a ) );
            }
        }
    }
}
template <class WorkTag, class FunctorType, class ListType, class ReduceType, bool Team>
__global__ void k_second_neighbor_reduce( FunctorType functor, ListType list, int begin,
                                          int end, ReduceType* result /* per-CTA partials */ )
{
    using traits = NeighborList<ListType>;
    ReduceType local = ReduceType();
    const int lane = threadIdx.x & 31;
    const int first = Team ? ( ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5 )
                           : ( blockIdx.x * blockDim.x + threadIdx.x );
    const int stride = Team ? ( ( gridDim.x * blockDim.x ) >> 5 ) : ( gridDim.x * blockDim.x );
    for ( int i = begin + first; i < end; i += stride )
    {
        const int nn = (int)traits::numNeighbor( list, i );
        for ( int n = Team ? lane : 0; n < nn; n += Team ? 32 : 1 )
        {
            const int j = (int)traits::getNeighbor( list, i, n );
            for ( int a = n + 1; a < nn; ++a )
                functorTagDispatch<WorkTag>( functor, i, j,
                                             (int)traits::getNeighbor( list, i, a ), local );
        }
    }
    block_reduce_store( local, result );
}
// neighbor_parallel_for directly on a LinkedCellList (LinkedCellParallelFor,
// Cabana_Parallel.hpp:1122-1290): every particle j != i of the stencil cells of i's bin is
// handed to the functor, which applies its own cutoff.  TEAM: one warp per particle.
template <class WorkTag, class FunctorType, bool TEAM>
__global__ void k_linked_cell_for( FunctorType functor, cb_lcl_view l, int begin, int end )
{
    const int lane = threadIdx.x & 31;
    const int first = TEAM ? ( ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5 )
                           : ( blockIdx.x * blockDim.x + threadIdx.x );
    const int stride = TEAM ? ( ( gridDim.x * blockDim.x ) >> 5 ) : ( gridDim.x * blockDim.x );
    const int sny = l.stencil_grid.nx[1], snz = l.stencil_grid.nx[2];
    const int ny = l.grid.nx[1], nz = l.grid.nx[2];
    for ( int i = begin + first; i < end; i += stride )
    {
        // getStencilCells( getParticleBin( i ) ) (Cabana_LinkedCellList.hpp:105-119, :841-858)
        const int cell = l.particle_bins[i - l.begin];
        const int ci = cell / ( sny * snz ), cj = ( cell / snz ) % sny, ck = cell % snz;
        const int R = l.cell_range;
        const int imin = ci - R > 0 ? ci - R : 0;
        const int imax = ci + R + 1 < l.stencil_grid.nx[0] ? ci + R + 1 : l.stencil_grid.nx[0];
        const int jmin = cj - R > 0 ? cj - R : 0;
        const int jmax = cj + R + 1 < sny ? cj + R + 1 : sny;
        const int kmin = ck - R > 0 ? ck - R : 0;
        const int kmax = ck + R + 1 < snz ? ck + R + 1 : snz;
        for ( int gi = imin; gi < imax; ++gi )
            for ( int gj = jmin; gj < jmax; ++gj )
                for ( int gk = kmin; gk < kmax; ++gk )
                {
                    const int c = ( gi * ny + gj ) * nz + gk; // cardinalBinIndex (:226-231)
                    const unsigned n0 = l.offsets[c];
                    const unsigned n1 = n0 + (unsigned)l.counts[c];
                    for ( unsigned n = n0 + ( TEAM ? lane : 0 ); n < n1; n += TEAM ? 32 : 1 )
                    {
                        // getParticle (:863-872)
                        const int j = l.sorted ? (int)( n + l.begin ) : (int)l.permute[n];
                        if ( j != i ) // NeighborDiscriminator<SelfNeighborTag>
                            functorTagDispatch<WorkTag>( functor, i, j );
                    }
                }
    }
}
// neighbor_parallel_reduce directly on a LinkedCellList (LinkedCellParallelReduce,
// Cabana_Parallel.hpp:1296-1468): functor( i, j, ival ) for every j != i of the stencil cells.
template <class WorkTag, class FunctorType, class ReduceType, bool TEAM>
__global__ void k_linked_cell_reduce( FunctorType functor, cb_lcl_view l, int begin, int end,
                                      ReduceType* partials )
{
    ReduceType local = ReduceType();
    const int lane = threadIdx.x & 31;
    const int first = TEAM ? ( ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5 )
                           : ( blockIdx.x * blockDim.x + threadIdx.x );
    const int stride = TEAM ? ( ( gridDim.x * blockDim.x ) >> 5 ) : ( gridDim.x * blockDim.x );
    const int sny = l.stencil_grid.nx[1], snz = l.stencil_grid.nx[2];
    const int ny = l.grid.nx[1], nz = l.grid.nx[2];
    for ( int i = begin + first; i < end; i += stride )
    {
        const int cell = l.particle_bins[i - l.begin];
        const int ci = cell / ( sny * snz ), cj = ( cell / snz ) % sny, ck = cell % snz;
        const int R = l.cell_range;
        const int imin = ci - R > 0 ? ci - R : 0;
        const int imax = ci + R + 1 < l.stencil_grid.nx[0] ? ci + R + 1 : l.stencil_grid.nx[0];
        const int jmin = cj - R > 0 ? cj - R : 0;
        const int jmax = cj + R + 1 < sny ? cj + R + 1 : sny;
        const int kmin = ck - R > 0 ? ck - R : 0;
        const int kmax = ck + R + 1 < snz ? ck + R + 1 : snz;
        for ( int gi = imin; gi < imax; ++gi )
            for ( int gj = jmin; gj < jmax; ++gj )
                for ( int gk = kmin; gk < kmax; ++gk )
                {
                    const int c = ( gi * ny + gj ) * nz + gk;
                    const unsigned n0 = l.offsets[c];
                    const unsigned n1 = n0 + (unsigned)l.counts[c];
                    for ( unsigned n = n0 + ( TEAM ? lane : 0 ); n < n1; n += TEAM ? 32 : 1 )
                    {
                        const int j = l.sorted ? (int)( n + l.begin ) : (int)l.permute[n];
                        if ( j != i )
                            functorTagDispatch<WorkTag>( functor, i, j, local );
                    }
                }
    }
    block_reduce_store( local, partials );
}
inline int grid_for( long long items, int block )
{
    long long b = ( items + block - 1 ) / block;
    if ( b > 148 * 32 )
        b = 148 * 32;
    return (int)( b < 1 ? 1 : b );
}
inline cudaStream_t stream_of( const DeviceExecutionSpace& space )
{
    return static_cast<cudaStream_t>( space.stream() );
}
inline void check_launch( const char* where )
{
    if ( cudaGetLastError() != cudaSuccess )
        throw std::runtime_error( std::string( where ) + ": launch failed" );
}
// Runs `launch( grid, block, smem_bytes, partials )` (a reduce kernel that leaves one partial
// per CTA), sums the partials on the device and returns the value -- the synchronisation a
// Kokkos::parallel_reduce into a host scalar implies.
template <class ReduceType, class Launch>
ReduceType run_reduce( const DeviceExecutionSpace& space, long long threads, Launch&& launch )
{
    ScopedRegion region( "Cabana::neighbor_parallel_reduce" ); // Cabana_Parallel.hpp:648
    constexpr int block = 256;
    const int grid = grid_for( threads, block );
    auto dev = device_alloc<ReduceType>( (std::size_t)grid + 1 );
    cudaStream_t st = stream_of( space );
    launch( grid, block, sizeof( ReduceType ) * block, dev.get() + 1 );
    check_launch( "Cabana::neighbor_parallel_reduce" );
    k_sum_partials<ReduceType>
        <<<1, block, sizeof( ReduceType ) * block, st>>>( dev.get() + 1, grid, dev.get() );
    check_launch( "Cabana::neighbor_parallel_reduce" );
    ReduceType out;
    check( cb_memcpy_d2h( &out, dev.get(), sizeof( ReduceType ), space.stream() ),
           "Cabana::neighbor_parallel_reduce" );
    check( cb_stream_synchronize( space.stream() ), "Cabana::neighbor_parallel_reduce" );
    return out;
}
} // namespace Impl

//! neighbor_parallel_for, FirstNeighborsTag x SerialOpTag (Cabana_Parallel.hpp:251-293)
template <class FunctorType, class NeighborListType, class WorkTag>
inline void neighbor_parallel_for( const RangePolicy<WorkTag>& exec_policy,
                                   const FunctorType& functor, const NeighborListType& list,
                                   const FirstNeighborsTag, const SerialOpTag,
                                   const std::string& = "",
                                   typename std::enable_if<
                                       !is_linked_cell_list<NeighborListType>::value,
                                       int>::type* = 0 )
{
    Impl::ScopedRegion region( "Cabana::neighbor_parallel_for" ); // Cabana_Parallel.hpp:259
    const int b = (int)exec_policy.begin(), e = (int)exec_policy.end();
    if ( e <= b )
        return;
    Impl::k_neighbor_for_serial<WorkTag>
        <<<Impl::grid_for( e - b, 256 ), 256, 0, Impl::stream_of( exec_policy.space() )>>>(
            functor, list.deviceView(), b, e );
    Impl::check_launch( "Cabana::neighbor_parallel_for" );
}
//! neighbor_parallel_for, FirstNeighborsTag x TeamOpTag (:386-435)
template <class FunctorType, class NeighborListType, class WorkTag>
inline void neighbor_parallel_for( const RangePolicy<WorkTag>& exec_policy,
                                   const FunctorType& functor, const NeighborListType& list,
                                   const FirstNeighborsTag, const TeamOpTag,
                                   const std::string& = "",
                                   typename std::enable_if<
                                       !is_linked_cell_list<NeighborListType>::value,
                                       int>::type* = 0 )
{
    Impl::ScopedRegion region( "Cabana::neighbor_parallel_for" ); // :394
    const int b = (int)exec_policy.begin(), e = (int)exec_policy.end();
    if ( e <= b )
        return;
    Impl::k_neighbor_for_team<WorkTag>
        <<<Impl::grid_for( (long long)( e - b ) * 32, 256 ), 256, 0,
           Impl::stream_of( exec_policy.space() )>>>( functor, list.deviceView(), b, e );
    Impl::check_launch( "Cabana::neighbor_parallel_for" );
}
//! neighbor_parallel_for on a LinkedCellList, no stored list (:1511-1595): Serial / Team.
//! Positions must be in the order the list currently describes (permuted iff sorted()).
template <class FunctorType, class M, class S, std::size_t D, class WorkTag, class OpTag>
inline void neighbor_parallel_for( const RangePolicy<WorkTag>& exec_policy,
                                   const FunctorType& functor,
                                   const LinkedCellList<M, S, D>& list, const FirstNeighborsTag,
                                   const OpTag, const std::string& = "" )
{
    static_assert( std::is_same<OpTag, SerialOpTag>::value ||
                       std::is_same<OpTag, TeamOpTag>::value,
                   "LinkedCellList traversal is Serial or Team" );
    const int b = (int)exec_policy.begin(), e = (int)exec_policy.end();
    if ( e <= b )
        return;
    if ( b < (int)list.getParticleBegin() || e > (int)list.getParticleEnd() )
        throw std::runtime_error( "Cabana::neighbor_parallel_for: range outside the binned range" );
    constexpr bool team = std::is_same<OpTag, TeamOpTag>::value;
    const long long threads = team ? (long long)( e - b ) * 32 : ( e - b );
    Impl::k_linked_cell_for<WorkTag, FunctorType, team>
        <<<Impl::grid_for( threads, 256 ), 256, 0, Impl::stream_of( exec_policy.space() )>>>(
            functor, list.deviceView().v, b, e );
    Impl::check_launch( "Cabana::neighbor_parallel_for" );
}
//! neighbor_parallel_reduce on a LinkedCellList (:1596-1726): Serial / Team.  The range must be
//! the binned range (the reference asserts equality, :1651-1652).
template <class FunctorType, class M, class S, std::size_t D, class ReduceType, class WorkTag,
          class OpTag>
inline void neighbor_parallel_reduce( const RangePolicy<WorkTag>& exec_policy,
                                      const FunctorType& functor,
                                      const LinkedCellList<M, S, D>& list,
                                      const FirstNeighborsTag, const OpTag,
                                      ReduceType& reduce_val, const std::string& = "" )
{
    static_assert( std::is_same<OpTag, SerialOpTag>::value ||
                       std::is_same<OpTag, TeamOpTag>::value,
                   "LinkedCellList traversal is Serial or Team" );
    const int b = (int)exec_policy.begin(), e = (int)exec_policy.end();
    if ( b != (int)list.getParticleBegin() || e != (int)list.getParticleEnd() )
        throw std::runtime_error(
            "Cabana::neighbor_parallel_reduce: cannot iterate over a range that was not binned" );
    constexpr bool team = std::is_same<OpTag, TeamOpTag>::value;
    const long long threads = team ? (long long)( e - b ) * 32 : ( e - b );
    const cb_lcl_view v = list.deviceView().v;
    cudaStream_t st = Impl::stream_of( exec_policy.space() );
    reduce_val = Impl::run_reduce<ReduceType>(
        exec_policy.space(), threads > 0 ? threads : 1,
        [&]( int grid, int block, std::size_t smem, ReduceType* partials )
        {
            Impl::k_linked_cell_reduce<WorkTag, FunctorType, ReduceType, team>
                <<<grid, block, smem, st>>>( functor, v, b, e, partials );
        } );
}
//! neighbor_parallel_for, SecondNeighborsTag x {Serial,Team,TeamVector} (:315-365, :458-594)
template <class FunctorType, class NeighborListType, class WorkTag, class OpTag>
inline void neighbor_parallel_for( const RangePolicy<WorkTag>& exec_policy,
                                   const FunctorType& functor, const NeighborListType& list,
                                   const SecondNeighborsTag, const OpTag,
                                   const std::string& = "" )
{
    const int b = (int)exec_policy.begin(), e = (int)exec_policy.end();
    if ( e <= b )
        return;
    using view_type = typename NeighborListType::device_view_type;
    cudaStream_t st = Impl::stream_of( exec_policy.space() );
    if ( std::is_same<OpTag, SerialOpTag>::value )
        Impl::k_second_neighbor_for_serial<WorkTag, FunctorType, view_type>
            <<<Impl::grid_for( e - b, 128 ), 128, 0, st>>>( functor, list.deviceView(), b, e );
    else if ( std::is_same<OpTag, TeamOpTag>::value )
        Impl::k_second_neighbor_for_team<WorkTag, FunctorType, view_type, false>
            <<<Impl::grid_for( (long long)( e - b ) * 32, 256 ), 256, 0, st>>>(
                functor, list.deviceView(), b, e );
    else
        Impl::k_second_neighbor_for_team<WorkTag, FunctorType, view_type, true>
            <<<Impl::grid_for( (long long)( e - b ) * 32, 256 ), 256, 0, st>>>(
                functor, list.deviceView(), b, e );
    Impl::check_launch( "Cabana::neighbor_parallel_for" );
}
//! neighbor_parallel_reduce, SecondNeighborsTag x {Serial,Team,TeamVector} (:704-759, :866-1001)
template <class FunctorType, class NeighborListType, class ReduceType, class WorkTag,
          class OpTag>
inline void neighbor_parallel_reduce( const RangePolicy<WorkTag>& exec_policy,
                                      const FunctorType& functor,
                                      const NeighborListType& list, const SecondNeighborsTag,
                                      const OpTag, ReduceType& reduce_val,
                                      const std::string& = "" )
{
    const int b = (int)exec_policy.begin(), e = (int)exec_policy.end();
    constexpr bool team = !std::is_same<OpTag, SerialOpTag>::value;
    const long long threads = team ? (long long)( e - b ) * 32 : ( e - b );
    auto dv = list.deviceView();
    cudaStream_t st = Impl::stream_of( exec_policy.space() );
    reduce_val = Impl::run_reduce<ReduceType>(
        exec_policy.space(), threads > 0 ? threads : 1,
        [&]( int grid, int block, std::size_t smem, ReduceType* partials )
        {
            Impl::k_second_neighbor_reduce<WorkTag, FunctorType,
                                           typename NeighborListType::device_view_type,
                                           ReduceType, team>
                <<<grid, block, smem, st>>>( functor, dv, b, e, partials );
        } );
}
//! neighbor_parallel_reduce, FirstNeighborsTag x {SerialOpTag,TeamOpTag} (:638-685, :787-844)
template <class FunctorType, class NeighborListType, class ReduceType, class WorkTag,
          class OpTag>
inline typename std::enable_if<!is_linked_cell_list<NeighborListType>::value>::type
neighbor_parallel_reduce( const RangePolicy<WorkTag>& exec_policy, const FunctorType& functor,
                          const NeighborListType& list, const FirstNeighborsTag, const OpTag,
                          ReduceType& reduce_val, const std::string& = "" )
{
    static_assert( std::is_same<OpTag, SerialOpTag>::value ||
                       std::is_same<OpTag, TeamOpTag>::value,
                   "first-neighbour reduce is Serial or Team" );
    const int b = (int)exec_policy.begin(), e = (int)exec_policy.end();
    constexpr bool team = std::is_same<OpTag, TeamOpTag>::value;
    const long long threads = team ? (long long)( e - b ) * 32 : ( e - b );
    auto dv = list.deviceView();
    cudaStream_t st = Impl::stream_of( exec_policy.space() );
    // Kokkos::parallel_reduce overwrites the result argument
    reduce_val = Impl::run_reduce<ReduceType>(
        exec_policy.space(), threads > 0 ? threads : 1,
        [&]( int grid, int block, std::size_t smem, ReduceType* partials )
        {
            Impl::k_neighbor_reduce<WorkTag, FunctorType,
                                    typename NeighborListType::device_view_type, ReduceType,
                                    team>
                <<<grid, block, smem, st>>>( functor, dv, b, e, partials );
        } );
}
//! for_each_neighbor inside a user kernel, thread-serial (:1058-1072)
template <class IndexType, class FunctorType, class NeighborListType>
CABANA_B200_DEVICE void for_each_neighbor( const IndexType i, const FunctorType& functor,
                                           const NeighborListType& list,
                                           const FirstNeighborsTag )
{
    using traits = NeighborList<NeighborListType>;
    for ( IndexType n = 0; n < (IndexType)traits::numNeighbor( list, i ); ++n )
        functor( i, (IndexType)traits::getNeighbor( list, i, n ) );
}
//! Team handle for the team form of for_each_neighbor: the warp of the calling thread
//! (stands in for Kokkos::TeamPolicy<>::member_type; league_rank() = global warp index).
struct WarpTeamMember
{
    CABANA_B200_DEVICE int team_rank() const { return (int)( threadIdx.x & 31u ); }
    CABANA_B200_DEVICE int team_size() const { return 32; }
    CABANA_B200_DEVICE int league_rank() const
    {
        return (int)( ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5 );
    }
    CABANA_B200_DEVICE int league_size() const { return (int)( ( gridDim.x * blockDim.x ) >> 5 ); }
    CABANA_B200_DEVICE void team_barrier() const { __syncwarp(); }
};
//! for_each_neighbor inside a user kernel, team-parallel over the neighbours (:1091-1110):
//! every thread of the team calls it with the same i; the neighbours are strided over the team
//! (Kokkos::TeamThreadRange).
template <class IndexType, class FunctorType, class NeighborListType, class TeamMemberType>
CABANA_B200_DEVICE void for_each_neighbor( const IndexType i, const TeamMemberType team,
                                           const FunctorType& functor,
                                           const NeighborListType& list,
                                           const FirstNeighborsTag )
{
    using traits = NeighborList<NeighborListType>;
    const IndexType nn = (IndexType)traits::numNeighbor( list, i );
    for ( IndexType n = (IndexType)team.team_rank(); n < nn; n += (IndexType)team.team_size() )
        Impl::functorTagDispatch<void>( functor, i, (IndexType)traits::getNeighbor( list, i, n ) );
}
#endif // __CUDACC__

} // namespace Cabana

#endif // CABANA_B200_HPP
