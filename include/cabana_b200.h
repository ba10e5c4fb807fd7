/* =============================================================================
 * cabana_b200.h -- C ABI of the B200-native neighbour-list path.
 *
 * This is the drop-in boundary for ONE hot path of ECP-copa/Cabana:
 *   LinkedCellList binning + permute  ->  VerletList build  ->
 *   neighbor_parallel_for / neighbor_parallel_reduce (Lennard-Jones consumer),
 *   plus the slab Halo / Distributor pack-unpack kernels.
 *
 * Cabana has no FFI: it is C++ templates over Kokkos.  Each entry point below
 * cites the reference interface (file:line under /root/reference) whose device
 * work it replaces; the header-only C++ shim in include/Cabana_B200.hpp
 * reproduces the reference's class/trait surface on top of these calls, and
 * INTEGRATION.md shows the binding a Cabana maintainer would add.
 *
 * Conventions
 *   - plain C types only; no torch / Kokkos types cross this boundary
 *   - every pointer is a DEVICE pointer unless the name ends in _h
 *   - all work is enqueued on the caller's cudaStream_t (`stream`, may be 0);
 *     calls that must size an allocation synchronise that stream once
 *   - returns cb_status (0 = ok); never throws; no CPU fallback -- without a
 *     CUDA device every compute entry point returns CB_ERR_CUDA
 *   - one caller thread per handle; handles own their device memory
 * ========================================================================== */
#ifndef CABANA_B200_H
#define CABANA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_VERSION_MAJOR 0
#define CB_VERSION_MINOR 1

typedef void* cb_stream_t; /* cudaStream_t */

typedef enum
{
    CB_OK = 0,
    CB_ERR_INVALID = 1,     /* bad argument (mirrors the reference's asserts) */
    CB_ERR_CUDA = 2,        /* CUDA runtime error; see cb_last_error_string() */
    CB_ERR_OVERFLOW = 3,    /* total neighbours exceed INT_MAX (reference: silent, Cabana_VerletList.hpp:518) */
    CB_ERR_UNSUPPORTED = 4, /* outside the hot path (2-D space, float positions, ...) */
    CB_ERR_NOMEM = 5
} cb_status;

/* Cabana::FullNeighborTag / HalfNeighborTag  (core/src/Cabana_NeighborList.hpp:35-50) */
typedef enum
{
    CB_NEIGHBOR_FULL = 0,
    CB_NEIGHBOR_HALF = 1
} cb_algorithm;

/* Cabana::VerletLayoutCSR / VerletLayout2D  (core/src/Cabana_VerletList.hpp:33-41) */
typedef enum
{
    CB_LAYOUT_CSR = 0,
    CB_LAYOUT_2D = 1
} cb_layout;

/* Cabana::SerialOpTag / TeamOpTag / TeamVectorOpTag  (core/src/Cabana_Parallel.hpp:199-212) */
typedef enum
{
    CB_OP_SERIAL = 0,
    CB_OP_TEAM = 1,
    CB_OP_TEAM_VECTOR = 2
} cb_op;

/* -----------------------------------------------------------------------------
 * Position field: a Cabana slice or a rank-2 Kokkos::View, described in place.
 *   element(i,d) = base[outer_stride*(i / vlen) + (i % vlen) + comp_stride*d]
 * Slice<double[3]>: outer_stride = Stride, vlen = VectorLength, comp_stride = VectorLength
 *   (core/src/Cabana_Slice.hpp:134-140; impl/Cabana_Index.hpp:62-80;
 *    Stride = sizeof(SoA)/sizeof(double), core/src/Cabana_AoSoA.hpp:192-196)
 * View<double**>(n,3): vlen = 1, outer_stride = stride(0), comp_stride = stride(1)
 * -------------------------------------------------------------------------- */
typedef struct
{
    const double* base;
    int64_t n;
    int64_t outer_stride;
    int32_t vlen;
    int64_t comp_stride;
} cb_positions;

/* A mutable per-particle field in the same layout (forces, or any member to permute /
 * gather / migrate).  elem_bytes is 4 or 8; num_comp is the product of the member's
 * trailing extents (core/src/Cabana_Sort.hpp:617-620). */
typedef struct
{
    void* base;
    int64_t n;
    int64_t outer_stride; /* in elements */
    int32_t vlen;
    int64_t comp_stride;  /* in elements */
    int32_t num_comp;
    int32_t elem_bytes;
} cb_field;

/* -----------------------------------------------------------------------------
 * Library / device
 * -------------------------------------------------------------------------- */
int cb_version(void);                     /* major*1000 + minor */
const char* cb_last_error_string(void);   /* thread-local, never NULL */
int cb_device_count(void);                /* 0 when no CUDA device is visible */
int cb_set_device(int device);
/* Number of CUDA kernels this library has launched in this process (evidence that the
 * CUDA path, not a fallback, did the work; bench.py reports the delta per timed region). */
int64_t cb_kernel_launch_count(void);

/* Device memory helpers for callers without a CUDA runtime of their own (the C++
 * shim and the Python mirror).  Plain cudaMalloc / cudaMemcpyAsync wrappers. */
int cb_malloc(void** ptr, size_t bytes);
int cb_free(void* ptr);
int cb_malloc_host(void** ptr_h, size_t bytes); /* pinned */
int cb_free_host(void* ptr_h);
int cb_memcpy_h2d(void* dst, const void* src_h, size_t bytes, cb_stream_t stream);
int cb_memcpy_d2h(void* dst_h, const void* src, size_t bytes, cb_stream_t stream);
int cb_memcpy_d2d(void* dst, const void* src, size_t bytes, cb_stream_t stream);
int cb_memset(void* dst, int value, size_t bytes, cb_stream_t stream);
int cb_stream_synchronize(cb_stream_t stream);

/* -----------------------------------------------------------------------------
 * CartesianGrid  (core/src/impl/Cabana_CartesianGrid.hpp:44-283) -- host helpers
 * with the reference's exact arithmetic (no contraction); the kernels use the same
 * inline functions on the device.
 * -------------------------------------------------------------------------- */
typedef struct
{
    double min[3];
    double max[3];
    double dx[3];
    double rdx[3];
    int32_t nx[3];
} cb_grid;

int cb_grid_init(cb_grid* grid, const double* min_h, const double* max_h,
                 const double* delta_h);                         /* :61-73 */
int cb_grid_locate_point(const cb_grid* grid, const double* p_h, int32_t* ijk_h); /* :171-182 */
double cb_grid_min_distance_to_point(const cb_grid* grid, const double* x_h,
                                     const int32_t* ijk_h);       /* :207-223 */
int cb_grid_cardinal_cell_index(const cb_grid* grid, int i, int j, int k); /* :226-231 */
int cb_grid_ijk_bin_index(const cb_grid* grid, int cardinal, int32_t* ijk_h); /* :251-258 */

/* LinkedCellStencil::getCells  (core/src/Cabana_LinkedCellList.hpp:55-66, :105-119) */
int cb_stencil_cell_range(double cell_size_ratio);
int cb_stencil_get_cells(const cb_grid* stencil_grid, int cell_range, int cell,
                         int32_t* min_h, int32_t* max_h);

/* -----------------------------------------------------------------------------
 * LinkedCellList  (core/src/Cabana_LinkedCellList.hpp:128-909)
 *
 * cb_lcl_create  <- constructors :166-460 (grid from delta/min/max; stencil from
 *                   radius/ratio; radius <= 0 selects the "simple" constructor's
 *                   stencil: radius = delta[0], ratio = 1, :186)
 * cb_lcl_build   <- build() :651-739 + storeParticleBins() :778-823
 *                   kernels: warp-aggregated atomic cell histogram + slot claim,
 *                   decoupled look-back exclusive scan, permute fill
 * cb_lcl_permute <- permute(LinkedCellList&, aosoa|slice|view) :1130-1145 and
 *                   permute(BinningData, slice) core/src/Cabana_Sort.hpp:600-656;
 *                   sets sorted = true and recomputes particle_bins
 * -------------------------------------------------------------------------- */
typedef struct cb_lcl cb_lcl;

typedef struct
{
    cb_grid grid;          /* _grid */
    cb_grid stencil_grid;  /* _cell_stencil.grid */
    int32_t cell_range;    /* _cell_stencil.cell_range */
    int32_t sorted;        /* _sorted */
    int64_t begin, end;    /* binned particle range */
    int64_t num_cells;     /* totalBins() */
    const int32_t* counts;        /* [num_cells]   binSize  */
    const uint32_t* offsets;      /* [num_cells+1] binOffset (size_type = unsigned on Kokkos CUDA) */
    const uint32_t* permute;      /* [end-begin]   permutation(): absolute particle ids */
    const int32_t* particle_bins; /* [end-begin]   getParticleBin(p) = particle_bins[p-begin] */
} cb_lcl_view;

int cb_lcl_create(cb_lcl** out, const double* delta_h, const double* min_h,
                  const double* max_h, double neighborhood_radius,
                  double cell_size_ratio);
int cb_lcl_build(cb_lcl* lcl, const cb_positions* x, int64_t begin, int64_t end,
                 cb_stream_t stream);
int cb_lcl_get(const cb_lcl* lcl, cb_lcl_view* view_h);
/* permute(BinningData, aosoa|slice|view) (core/src/Cabana_Sort.hpp:549-715): for each field
 * tmp[i] = data[permute[i]], data[begin+i] = tmp[i], i in [0, end-begin).  `permute` is a device
 * array of absolute particle ids (BinningData::permutation, :96-100). */
int cb_binning_permute(int64_t begin, int64_t end, const uint32_t* permute,
                       const cb_field* fields_h, int num_fields, cb_stream_t stream);
int cb_lcl_permute(cb_lcl* lcl, const cb_field* fields_h, int num_fields,
                   cb_stream_t stream);
int cb_lcl_update(cb_lcl* lcl, int sorted); /* update(bool) :829 */
int cb_lcl_destroy(cb_lcl* lcl);

/* -----------------------------------------------------------------------------
 * VerletList  (core/src/Cabana_VerletList.hpp:824-1495)
 *
 * cb_verlet_build <- VerletList::build(ExecutionSpace, x, begin, end, radius,
 *                    cell_size_ratio, grid_min, grid_max, max_neigh) :1351-1392
 *                    = VerletListBuilder::init :205-240 (bins ALL particles on the
 *                    delta = ratio*radius grid) + buildImpl :1441-1484
 *                    (count :316-474, processCounts :493-562, fill :572-713).
 * Results (cb_verlet_get) = VerletListData<CSR|2D> :50-113:
 *   CSR: counts[n], offsets[n] (exclusive scan of counts in particle order),
 *        neighbors[total]
 *   2D : counts[n], neighbors(i,k) at neighbors[i*row_stride + k*col_stride],
 *        extent(1) = width, max_n = true maximum (:94-96)
 * Per-particle neighbour SETS are bit-identical to the reference's; the order of
 * ids inside a row is unspecified there (atomic slot claim, :69-70) and is
 * deterministic here (stencil order x binned order).
 * A handle may be rebuilt; device buffers are reused when large enough.
 * -------------------------------------------------------------------------- */
typedef struct cb_verlet cb_verlet;

typedef struct
{
    int32_t layout;     /* cb_layout */
    int32_t algorithm;  /* cb_algorithm */
    int64_t n;          /* counts.extent(0) = size(x) */
    int32_t* counts;
    int32_t* offsets;   /* CSR only, else NULL */
    int32_t* neighbors;
    int64_t total;      /* NeighborList::totalNeighbor */
    int64_t max_n;      /* NeighborList::maxNeighbor  */
    int64_t width;      /* 2D: neighbors.extent(1); CSR: 0 */
    int64_t row_stride; /* 2D element (i,k) = neighbors[i*row_stride + k*col_stride] */
    int64_t col_stride;
    int32_t refilled;   /* 2D with max_neigh: 1 if the realloc+refill path ran (:554-561) */
    int64_t extent;     /* neighbors.extent(0) (CSR): == total with reference row placement,
                           >= total with CB_ROWS_BINNED (rows are not packed) */
} cb_verlet_view;

int cb_verlet_create(cb_verlet** out);
int cb_verlet_build(cb_verlet* list, const cb_positions* x, int64_t begin,
                    int64_t end, double neighborhood_radius,
                    double cell_size_ratio, const double* grid_min_h,
                    const double* grid_max_h, int64_t max_neigh, int algorithm,
                    int layout, int build_op, cb_stream_t stream);
int cb_verlet_get(const cb_verlet* list, cb_verlet_view* view_h);
/* Where the CSR rows live inside `neighbors` (set before cb_verlet_build; 2D lists ignore it).
 *   CB_ROWS_REFERENCE (default): offsets = exclusive scan of counts in particle order, rows
 *       packed, exactly the reference's arrays (Cabana_VerletList.hpp:478-491, :507-532).
 *   CB_ROWS_BINNED: rows stay where the single test pass wrote them (cell order, small gaps
 *       between warps' reservations); offsets[i] is still the start of row i, so everything
 *       that goes through NeighborList<>::getNeighbor / neighbor_parallel_for is unchanged
 *       (the reference leaves row placement to the implementation: rows are claimed with an
 *       atomic, :69-70), but offsets is not monotone and neighbors.extent(0) > total.  Saves
 *       the offsets scan and the row-reorder pass (~23 % of a build). */
#define CB_ROWS_REFERENCE 0
#define CB_ROWS_BINNED 1
int cb_verlet_set_row_placement(cb_verlet* list, int placement);
/* setNeighbor(particle, neighbor, new_index) :1488-1494 */
int cb_verlet_set_neighbor(cb_verlet* list, int64_t particle_index,
                           int64_t neighbor_index, int32_t new_index,
                           cb_stream_t stream);
int cb_verlet_destroy(cb_verlet* list);

/* Optional per-phase device timing (CUDA events on the caller's stream):
 * ms_h[0..5] = binning, gather-permute / tile plan, count pass, offsets scan + max/sum, fill
 * pass, whole build -- AVERAGED over the builds (the 32 most recent at most) since the last
 * cb_verlet_set_profiling(list, 1), so a rebuild loop is timed without a host wait per build;
 * cb_verlet_get_phase_times waits for the last one.  Used by bench.py for the per-kernel
 * roofline; off by default. */
#define CB_VERLET_NUM_PHASES 6
int cb_verlet_set_profiling(cb_verlet* list, int enable);
int cb_verlet_get_phase_times(const cb_verlet* list, double* ms_h);

/* Per-particle cutoff radius: VerletList( x, begin, end, background_radius, neighborhood_radius
 * (slice or view, one double per particle), cell_size_ratio, grid_min, grid_max, max_neigh )
 * (core/src/Cabana_VerletList.hpp:989-1017 -> VerletListBuilder :181-203; withinCutoff /
 * neighborNotWithinCutoff / addNeighbor :244-305).  The linked cells use background_radius;
 * row i holds every valid j with dist_sqr <= radius(i)^2, plus every j that found i with
 * dist_sqr >= radius(i)^2 ("j adds itself because i will not find it").  counts/offsets describe
 * exactly those rows: the reference's count pass books the extra entry on the other row
 * (SURVEY.md Appendix B.6); its post-fill counts, which its test checks, are these. */
int cb_verlet_build_radii(cb_verlet* list, const cb_positions* x, const cb_field* radii_h,
                          int64_t begin, int64_t end, double background_radius,
                          double cell_size_ratio, const double* grid_min_h,
                          const double* grid_max_h, int64_t max_neigh, int algorithm,
                          int layout, int build_op, cb_stream_t stream);

/* NumSpaceDim = 2 (VerletList<..., 2>; Cabana_VerletList.hpp:377-392, :626-639): x2 describes
 * (n,2) positions (element (i,d), d < 2); grid_min2 / grid_max2 hold two values.  The list is
 * built by the 3-D kernels on [x, y, 0] with one shared z cell, which reproduces the 2-D cell
 * prune, distances and half criterion bit for bit. */
int cb_verlet_build_2d(cb_verlet* list, const cb_positions* x2, int64_t begin, int64_t end,
                       double neighborhood_radius, double cell_size_ratio,
                       const double* grid_min2_h, const double* grid_max2_h, int64_t max_neigh,
                       int algorithm, int layout, int build_op, cb_stream_t stream);

/* Self-test of the tensor-core distance filter the build relies on (tests only): runs the
 * count pass over `x` with EVERY filter value compared with the exact FP64 arithmetic.
 * out_h[0] = largest |filter - exact| observed over pairs with s <= 4 r^2 (the range the bound
 * is proven for), out_h[1] = the proven bound the in/out decisions assume for one tile (values
 * closer than twice the bound to the cutoff go to the exact tier), out_h[2] = number of such
 * pairs whose error exceeded the bound, out_h[3] = number of values at ANY distance that lie
 * outside the exact tier's band and have the wrong sign (must be 0). out_h holds 4 doubles. */
int cb_verlet_filter_selftest(cb_verlet* list, const cb_positions* x,
                              double neighborhood_radius, const double* grid_min_h,
                              const double* grid_max_h, int algorithm, double* out_h,
                              cb_stream_t stream);

/* Host-buffer entry points (the end-to-end path): positions live in HOST memory
 * (x_h->base is a host pointer; pinned memory makes the copies asynchronous), the
 * list is built on the device and copied back into caller-provided host arrays.
 * neighbors_h must hold neighbors_capacity ints; CB_ERR_NOMEM if too small
 * (query with cb_verlet_get after cb_verlet_build_host). */
int cb_verlet_build_host(cb_verlet* list, const cb_positions* x_h, int64_t begin,
                         int64_t end, double neighborhood_radius,
                         double cell_size_ratio, const double* grid_min_h,
                         const double* grid_max_h, int64_t max_neigh,
                         int algorithm, int layout, int build_op,
                         cb_stream_t stream);
int cb_verlet_copy_to_host(const cb_verlet* list, int32_t* counts_h,
                           int32_t* offsets_h, int32_t* neighbors_h,
                           int64_t neighbors_capacity, cb_stream_t stream);

/* -----------------------------------------------------------------------------
 * neighbor_parallel_for / neighbor_parallel_reduce, FirstNeighborsTag
 * (core/src/Cabana_Parallel.hpp:251-293 Serial, :386-435 Team, :638-685 and
 *  :787-844 reduce).  User functors cannot cross a C ABI (the C++ shim
 * instantiates them as templates); these are the pre-compiled consumers.
 *
 * Lennard-Jones functor for pair (i,j), r2 = |x_i - x_j|^2 < rc^2:
 *     sr2 = sigma^2/r2; sr6 = sr2^3; fpair = 24 eps sr6 (2 sr6 - 1)/r2
 *     f_i += fpair (x_i - x_j);   newton != 0 (half lists): f_j -= fpair (x_i - x_j)
 * `f` is ACCUMULATED into (zero it first), as a Kokkos functor would.
 * op = CB_OP_SERIAL: one thread per particle; CB_OP_TEAM: one warp per particle.
 * -------------------------------------------------------------------------- */
int cb_neighbor_for_lj(const cb_verlet_view* list_h, const cb_positions* x,
                       const cb_field* f, double eps, double sigma, double rc,
                       int newton, int op, int64_t begin, int64_t end,
                       cb_stream_t stream);
/* energy_h receives scale * sum over stored pairs of 4 eps (sr6^2 - sr6); the
 * call synchronises the stream (a reduce returns its value to the host). */
int cb_neighbor_reduce_lj(const cb_verlet_view* list_h, const cb_positions* x,
                          double eps, double sigma, double rc, double scale,
                          int op, int64_t begin, int64_t end, double* energy_h,
                          cb_stream_t stream);
/* The reference unit tests' functor (core/unit_test/neighbor_unit_test.hpp:291-348):
 * result[i] += j for every stored neighbour j of i. */
int cb_neighbor_for_id_sum(const cb_verlet_view* list_h, int64_t* result, int op,
                           int64_t begin, int64_t end, cb_stream_t stream);

/* neighbor_parallel_for directly on a LinkedCellList, no stored list:
 * LinkedCellParallelFor (core/src/Cabana_Parallel.hpp:1122-1290, :1511-1595) through
 * NeighborList<LinkedCellList> (core/src/Cabana_LinkedCellList.hpp:1149-1303): for i in
 * [begin,end) every particle j != i of the stencil cells of i's bin is handed to the functor,
 * which applies its own cutoff.  x must be in the order the list currently describes
 * (permuted if cb_lcl_permute was called).  op = CB_OP_SERIAL | CB_OP_TEAM.
 *   cb_lcl_neighbor_for_lj:    f_i += LJ pair force for r2 < rc^2
 *   cb_lcl_neighbor_for_count: result[i] += 1 for r2 <= cutoff^2 (tstLinkedCellList.hpp:704-780) */
int cb_lcl_neighbor_for_lj(const cb_lcl* lcl, const cb_positions* x, const cb_field* f,
                           double eps, double sigma, double rc, int op, int64_t begin,
                           int64_t end, cb_stream_t stream);
int cb_lcl_neighbor_for_count(const cb_lcl* lcl, const cb_positions* x, double cutoff,
                              int32_t* result, int op, int64_t begin, int64_t end,
                              cb_stream_t stream);

/* -----------------------------------------------------------------------------
 * Slab Halo / Distributor kernels  (core/src/Cabana_CommunicationPlanBase.hpp:96-224
 * countSendsAndCreateSteering, :596-657 createExportSteering;
 * core/src/impl/Cabana_Halo_Mpi.hpp:58-65/:113-121 gather pack/unpack, :269-282/
 * :334-347 scatter; core/src/impl/Cabana_Migrate_Mpi.hpp:92-103/:164-172).
 * The exchange itself (MPI_Send/Irecv in the reference) is NCCL send/recv on the same
 * stream, issued by the host layer between pack and unpack.
 * -------------------------------------------------------------------------- */
/* Ghost selection for a 1-D slab decomposition along x (compare-to-local-corners logic
 * of grid/src/Cabana_Grid_ParticleDistributor.hpp:123-143).  For owned particle i:
 *   export_ranks[2i]   = lo_rank if lo_rank >= 0 and x <  lo_thresh, else -1
 *   export_ranks[2i+1] = hi_rank if hi_rank >= 0 and x >= hi_thresh, else -1
 *   export_ids[2i] = export_ids[2i+1] = i
 * i.e. the (export_ids, export_ranks) pair a Cabana::Halo is constructed from
 * (core/src/Cabana_Halo.hpp:107-115). */
int cb_slab_halo_select(const cb_positions* x, int64_t num_local, double lo_thresh,
                        double hi_thresh, int lo_rank, int hi_rank,
                        int32_t* export_ranks /* [2*num_local] */,
                        uint32_t* export_ids /* [2*num_local] */, cb_stream_t stream);
/* Fused slab-halo plan: selection + stable compaction in ONE kernel.  On return
 * (synchronises once) counts_h[0] / counts_h[1] = exports to the lower / upper neighbour and
 * steer_lo[0..counts_h[0]) / steer_hi[0..counts_h[1]) = their particle ids in ascending order
 * -- the same plan cb_slab_halo_select + cb_comm_count_and_steer produce. */
int cb_slab_halo_plan(const cb_positions* x, int64_t num_local, double lo_thresh,
                      double hi_thresh, int has_lo, int has_hi, uint32_t* steer_lo,
                      uint32_t* steer_hi, int64_t* counts_h, cb_stream_t stream);
/* Migration destination = slab g with bounds_h[g] <= x < bounds_h[g+1] (the last slab
 * owns its upper face); -1 (dropped, impl/Cabana_CommunicationPlan_Mpi.hpp:98-103)
 * outside [bounds_h[0], bounds_h[num_ranks]].  num_ranks <= 64. */
int cb_slab_migrate_destinations(const cb_positions* x, int64_t num_local,
                                 const double* bounds_h /* [num_ranks+1] */, int num_ranks,
                                 int32_t* export_ranks, cb_stream_t stream);
/* countSendsAndCreateSteering: export_ranks[k] in [-1, num_ranks); on return
 * (synchronises) counts_h[num_ranks], offsets_h[num_ranks+1] (exclusive prefix) and
 * steering[offsets_h[r] + s] = export_ids[k] (k itself when export_ids is NULL) for the
 * s-th export with destination r: ascending rank, ascending k inside a rank
 * (deterministic; the reference's order inside a block is not). num_ranks <= 1024.
 * Cost: one histogram pass, then a stable radix partition of the exports by destination --
 * ONE histogram / scan / scatter sweep when at most 32 destinations receive something (any
 * Cartesian halo or migration), two up to 1024 -- independent of the communicator's size. */
int cb_comm_count_and_steer(const int32_t* export_ranks, int64_t num_export,
                            int num_ranks, int64_t* counts_h, int64_t* offsets_h,
                            uint32_t* steering, const uint32_t* export_ids,
                            cb_stream_t stream);
/* send_buffer tuple j = (field_0(steering[j]), field_1(...), ...): members in the given
 * order, each aligned to its element size, tuple padded to 8 bytes (cb_comm_tuple_bytes). */
int cb_comm_pack(const cb_field* fields_h, int num_fields, const uint32_t* steering,
                 int64_t count, void* send_buffer, cb_stream_t stream);
/* field(dst_begin + j) = recv_buffer[j]  (gather unpack / migrate unpack) */
int cb_comm_unpack(const cb_field* fields_h, int num_fields, int64_t dst_begin,
                   int64_t count, const void* recv_buffer, cb_stream_t stream);
/* scatter: field(steering[j]) += recv_buffer[j]  (atomic add; doubles only: tuples of
 * num_comp doubles, no padding) */
int cb_comm_scatter_add(const cb_field* field_h, const uint32_t* steering,
                        int64_t count, const void* recv_buffer, cb_stream_t stream);
/* The same for every arithmetic slice value type Cabana::scatter accepts
 * (impl/Cabana_Halo_Mpi.hpp:334-347): recv_buffer holds `count` packed tuples of this ONE field
 * in the layout cb_comm_pack / cb_comm_pack_range write (cb_comm_tuple_bytes(field, 1) bytes per
 * tuple: num_comp values of `dtype`, padded to 8 bytes). */
enum { CB_DTYPE_F64 = 0, CB_DTYPE_F32 = 1, CB_DTYPE_I32 = 2, CB_DTYPE_I64 = 3 };
int cb_comm_scatter_add_typed(const cb_field* field_h, const uint32_t* steering,
                              int64_t count, const void* recv_buffer, int dtype,
                              cb_stream_t stream);
/* send_buffer tuple j = fields(src_begin + j): packs a contiguous range, e.g. the ghost block
 * a scatter sends home (impl/Cabana_Halo_Mpi.hpp:269-300) */
int cb_comm_pack_range(const cb_field* fields_h, int num_fields, int64_t src_begin,
                       int64_t count, void* send_buffer, cb_stream_t stream);
/* bytes of one packed tuple */
int64_t cb_comm_tuple_bytes(const cb_field* fields_h, int num_fields);

/* -----------------------------------------------------------------------------
 * Peer-memory halo for the slab decomposition (one process per GPU on one NVSwitch box).
 *
 * Instead of pack -> NCCL send/recv -> unpack with a host-side count exchange, each rank owns
 * a receive WINDOW in its HBM that its two slab neighbours map through CUDA IPC; the fused
 * push below selects the ghost layer, compacts it (stable) and packs it STRAIGHT INTO the
 * neighbours' windows over NVLink, then publishes (count, sequence) in the window header.
 * The receiver waits on its own headers, reads the two counts back (the only host sync of
 * the exchange) and unpacks from local memory.  Windows are double-buffered by sequence
 * parity, which is sufficient because a rank can run at most one exchange ahead of a
 * neighbour.  This replaces impl/Cabana_Halo_Mpi.hpp:70-124 (Irecv/Send/Waitall/Barrier)
 * and the count exchange of impl/Cabana_CommunicationPlan_Mpi.hpp:152-178 for this topology.
 * -------------------------------------------------------------------------- */
typedef struct cb_p2p_window cb_p2p_window;
#define CB_IPC_HANDLE_BYTES 64
/* capacity_tuples per buffer; tuple_bytes from cb_comm_tuple_bytes */
/* (capacity_tuples and tuple_bytes must be the same on both sides of a face: the pusher computes
 * the offsets inside the neighbour's window from the values IT passes to cb_slab_halo_push /
 * cb_slab_step.) */
int cb_p2p_window_create(cb_p2p_window** out, int64_t capacity_tuples, int64_t tuple_bytes);
int cb_p2p_window_destroy(cb_p2p_window* w);
int cb_p2p_window_get_handle(const cb_p2p_window* w, void* handle_h /* 64 bytes */);
int cb_p2p_window_open(const void* handle_h, void** peer_base);   /* in the PEER process */
/* The window's base address for a pusher in the SAME process (several slabs driven by one
 * process): pass it as peer_lo / peer_hi; never cb_p2p_window_close it. */
int cb_p2p_window_local_base(const cb_p2p_window* w, void** base);
int cb_p2p_window_close(void* peer_base);
/* Fused plan + pack into the neighbours' windows (peer_lo / peer_hi from cb_p2p_window_open,
 * NULL when there is no neighbour on that side).  steer_scratch holds 2*num_local ids.
 * No host synchronisation. */
int cb_slab_halo_push(const cb_positions* x, const cb_field* fields_h, int num_fields,
                      int64_t num_local, double lo_thresh, double hi_thresh, void* peer_lo,
                      void* peer_hi, int64_t capacity_tuples, uint64_t sequence,
                      uint32_t* steer_scratch, cb_stream_t stream);
/* Wait (on the device, bounded) for the pushes of `sequence` into my windows, return the ghost
 * counts and the local buffers holding them; synchronises the stream once.
 * CB_ERR_NOMEM if a neighbour had more ghosts than capacity, CB_ERR_CUDA on time-out. */
int cb_slab_halo_wait(cb_p2p_window* from_lo, cb_p2p_window* from_hi, uint64_t sequence,
                      int64_t* counts_h /* [2] */, const void** data_lo, const void** data_hi,
                      cb_stream_t stream);

/* The whole sharded step in one host call (bench.py, N > 1): ghost selection + push into the
 * neighbours' windows, then ONE kernel that waits for their pushes and unpacks both ghost layers
 * behind the owned particles (lower neighbour's first), then cb_verlet_build with begin = 0,
 * end = num_local on x_all (whose n is its CAPACITY: owned + room for ghosts; the fields must
 * have the same capacity).  The ghost counts stay on the device -- the build reads the particle
 * count there -- and reach the host with the build's single read-back, so the step has no host
 * synchronisation before the list size is known (CB_SLAB_SYNC=1 selects the older
 * push / cb_slab_halo_wait / cb_comm_unpack / cb_verlet_build sequence with a read-back of the
 * counts; same result).  counts_h[0..1] = ghosts received from the lower / upper neighbour.
 * Replaces Halo construction + gather (impl/Cabana_Halo_Mpi.hpp:41-125) followed by
 * VerletList::build (Cabana_VerletList.hpp:1351-1392). */
int cb_slab_step(cb_verlet* list, const cb_positions* x_all, const cb_field* fields_h,
                 int num_fields, int64_t num_local, double lo_thresh, double hi_thresh,
                 void* peer_lo, void* peer_hi, cb_p2p_window* from_lo, cb_p2p_window* from_hi,
                 int64_t capacity_tuples, uint64_t sequence, uint32_t* steer_scratch,
                 double neighborhood_radius, double cell_size_ratio, const double* grid_min_h,
                 const double* grid_max_h, int64_t max_neigh, int algorithm, int layout,
                 int build_op, int64_t* counts_h, cb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CABANA_B200_H */
