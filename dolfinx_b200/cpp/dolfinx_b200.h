// dolfinx_b200.h — header-only C++20 mirror of the DOLFINx classes on the assembly hot path,
// implemented over the C-ABI of libbfx.so (include/bfx.h).  Class and method names, argument
// meaning and error behaviour follow the reference (paths relative to /root/reference/cpp/dolfinx):
//
//   common::IndexMap      common/IndexMap.h:94-331        (accessor subset + known-src/dest ctor)
//   common::Scatterer     common/Scatterer.h:46-538
//   la::SparsityPattern   la/SparsityPattern.h:25-180
//   la::Vector            la/Vector.h:47-422, free functions :434-514
//   la::MatrixCSR         la/MatrixCSR.h:67-624
//   fem::Form / DofMap / FunctionSpace / Function / Constant / DirichletBC
//                         fem/Form.h:116-668, fem/DofMap.h, fem/FunctionSpace.h, fem/DirichletBC.h:262-601
//   fem::assemble_matrix / assemble_vector / apply_lifting / set_diagonal
//                         fem/assembler.h:230-257, 336-493, 513-630, 644-686
//
// Storage lives in HBM (DeviceArray); every arithmetic method forwards to a CUDA kernel through
// the C-ABI.  There is no MPI in this build environment: a rank is a process (or thread) with one GPU
// and a bfx_comm_t (NCCL) for the data path.  The HOST side of several ranks - IndexMap src/dest and
// offsets, the Scatterer plan (common/Scatterer.h:65-198), SparsityPattern::finalize with its ghost-row
// exchange (la/SparsityPattern.cpp:264-491) and the MatrixCSR ghost-row plan (la/MatrixCSR.h:705-849)
// - is implemented here on two exchange callbacks of common::Comm (neighbor_alltoallv, allgatherv: what
// MPI_Neighbor_alltoallv / MPI_Allgatherv do in the reference), so that MPI, a torch.distributed bridge
// or the in-process mailboxes of tests/cpp/test_cpp_multirank.cpp can be plugged.
#pragma once

#include "../../include/bfx.h"
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <numeric>
#include <optional>
#include <span>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <variant>
#include <vector>

namespace dolfinx_b200
{
/// Convert a C-ABI status into the exception the reference would throw
inline void check(int status)
{
  if (status == BFX_OK)
    return;
  if (status == BFX_ERR_NOT_IN_SPARSITY)
    throw std::runtime_error("Entry not in sparsity"); // la/matrix_csr_impl.h:93-94
  throw std::runtime_error(std::string(bfx_status_string(status)) + ": " + bfx_last_error());
}

/// RAII device buffer (the "Container" template argument of la::Vector / la::MatrixCSR,
/// la/Vector.h:42-49, la/MatrixCSR.h:67-70, with device storage)
template <typename T>
class DeviceArray
{
public:
  DeviceArray() = default;
  explicit DeviceArray(std::size_t n, bool zero = true) : _n(n)
  {
    void* p = nullptr;
    check(bfx_malloc(&p, n * sizeof(T)));
    _p = static_cast<T*>(p);
    if (zero and n)
      check(bfx_memset(_p, 0, n * sizeof(T), nullptr));
  }
  explicit DeviceArray(std::span<const T> host) : DeviceArray(host.size(), false) { copy_from(host); }
  DeviceArray(const DeviceArray&) = delete;
  DeviceArray& operator=(const DeviceArray&) = delete;
  DeviceArray(DeviceArray&& o) noexcept : _p(o._p), _n(o._n) { o._p = nullptr, o._n = 0; }
  DeviceArray& operator=(DeviceArray&& o) noexcept
  {
    std::swap(_p, o._p);
    std::swap(_n, o._n);
    return *this;
  }
  ~DeviceArray() { bfx_free(_p); }
  T* data() { return _p; }
  const T* data() const { return _p; }
  std::size_t size() const { return _n; }
  void copy_from(std::span<const T> host)
  {
    if (host.size() > _n)
      throw std::runtime_error("DeviceArray::copy_from: size mismatch");
    check(bfx_memcpy(_p, host.data(), host.size() * sizeof(T), nullptr));
    check(bfx_stream_sync(nullptr));
  }
  std::vector<T> to_host() const
  {
    std::vector<T> h(_n);
    check(bfx_memcpy(h.data(), _p, _n * sizeof(T), nullptr));
    check(bfx_stream_sync(nullptr));
    return h;
  }
  void fill_zero() { check(bfx_memset(_p, 0, _n * sizeof(T), nullptr)); }

private:
  T* _p = nullptr;
  std::size_t _n = 0;
};

// -------------------------------------------------------------------------------------------------
namespace common
{
/// The rank-local view of a communicator: rank, size and the NCCL handle of the data path
struct Comm
{
  int rank = 0, size = 1;
  bfx_comm_t* nccl = nullptr;
  /// Host-side exchanges of the plan constructors (integer data; needed when size > 1):
  /// neighbor_alltoallv(dest, src, send): send[i] goes to rank dest[i]; returns the buffer received from every rank
  /// src[j], in the order of src (MPI_Neighbor_alltoallv on the graph src -> me -> dest).
  using Buffers = std::vector<std::vector<std::int64_t>>;
  std::function<Buffers(std::span<const int>, std::span<const int>, const Buffers&)> neighbor_alltoallv;
  /// allgatherv(mine): the arrays of all ranks, by rank (MPI_Allgatherv)
  std::function<Buffers(std::span<const std::int64_t>)> allgatherv;
};

class IndexMap
{
public:
  /// IndexMap(comm, local_size) on one rank (common/IndexMap.cpp:865-886)
  IndexMap(Comm comm, std::int32_t local_size) : _comm(comm), _local_range{0, local_size}, _size_global(local_size)
  {
    if (comm.size != 1)
      throw std::runtime_error("IndexMap(comm, n): pass the global offset/size for several ranks");
  }
  /// IndexMap with known src/dest (common/IndexMap.cpp:899-932); the exclusive scan / all-reduce
  /// results (offset, global size) are passed in.
  IndexMap(Comm comm, std::int32_t local_size, std::int64_t offset, std::int64_t size_global,
           const std::array<std::vector<int>, 2>& src_dest, std::span<const std::int64_t> ghosts,
           std::span<const int> owners)
      : _comm(comm), _local_range{offset, offset + local_size}, _size_global(size_global),
        _ghosts(ghosts.begin(), ghosts.end()), _owners(owners.begin(), owners.end()), _src(src_dest[0]),
        _dest(src_dest[1])
  {
    if (ghosts.size() != owners.size() or !std::ranges::is_sorted(_src) or !std::ranges::is_sorted(_dest))
      throw std::runtime_error("IndexMap: inconsistent ghost data");
  }
  /// IndexMap(comm, local_size, ghosts, owners) on several ranks (common/IndexMap.cpp:888-932): collective.
  /// Offset by exclusive scan, global size by reduction, src = sorted unique ghost owners, dest = the ranks that
  /// list this rank among their src (build_src_dest).
  IndexMap(Comm comm, std::int32_t local_size, std::span<const std::int64_t> ghosts, std::span<const int> owners)
      : _comm(comm), _ghosts(ghosts.begin(), ghosts.end()), _owners(owners.begin(), owners.end())
  {
    if (ghosts.size() != owners.size())
      throw std::runtime_error("IndexMap: inconsistent ghost data");
    _src.assign(owners.begin(), owners.end());
    std::ranges::sort(_src);
    _src.erase(std::unique(_src.begin(), _src.end()), _src.end());
    std::int64_t offset = 0, total = local_size;
    if (comm.size > 1)
    {
      if (!comm.allgatherv)
        throw std::runtime_error("IndexMap on several ranks needs Comm::allgatherv");
      std::vector<std::int64_t> mine{local_size};
      mine.insert(mine.end(), _src.begin(), _src.end());
      const Comm::Buffers all = comm.allgatherv(mine);
      total = 0;
      for (int r = 0; r < comm.size; ++r)
      {
        if (r < comm.rank)
          offset += all[r][0];
        total += all[r][0];
        if (std::find(all[r].begin() + 1, all[r].end(), (std::int64_t)comm.rank) != all[r].end())
          _dest.push_back(r);
      }
    }
    _local_range = {offset, offset + local_size};
    _size_global = total;
  }
  std::array<std::int64_t, 2> local_range() const noexcept { return _local_range; }
  std::int32_t num_ghosts() const noexcept { return static_cast<std::int32_t>(_ghosts.size()); }
  std::int32_t size_local() const noexcept { return static_cast<std::int32_t>(_local_range[1] - _local_range[0]); }
  std::int64_t size_global() const noexcept { return _size_global; }
  std::span<const std::int64_t> ghosts() const noexcept { return _ghosts; }
  std::span<const int> owners() const noexcept { return _owners; }
  std::span<const int> src() const noexcept { return _src; }
  std::span<const int> dest() const noexcept { return _dest; }
  const Comm& comm() const { return _comm; }
  /// common/IndexMap.cpp:957-974
  void local_to_global(std::span<const std::int32_t> local, std::span<std::int64_t> global) const
  {
    const std::int32_t n = size_local();
    for (std::size_t i = 0; i < local.size(); ++i)
      global[i] = local[i] < n ? _local_range[0] + local[i] : _ghosts[local[i] - n];
  }

private:
  Comm _comm;
  std::array<std::int64_t, 2> _local_range;
  std::int64_t _size_global;
  std::vector<std::int64_t> _ghosts;
  std::vector<int> _owners, _src, _dest;
};

/// Plan + device exchange.  Members are the reference's (common/Scatterer.h:503-537).
class Scatterer
{
public:
  /// One rank: all plan arrays empty, every begin/end a no-op (common/Scatterer.h:71-72)
  /// Several ranks: the plan of common/Scatterer.h:84-197 (collective: one neighbourhood exchange of ghost indices)
  Scatterer(const IndexMap& map, int bs)
      : _src(map.src().begin(), map.src().end()), _dest(map.dest().begin(), map.dest().end()), _nccl(map.comm().nccl)
  {
    if (map.comm().size == 1)
      return;
    if (!map.comm().neighbor_alltoallv)
      throw std::runtime_error("Scatterer on several ranks needs Comm::neighbor_alltoallv");
    // ghost positions sorted by owner, stable (:98-101); sizes / displacements per owning rank (:122-130)
    const std::span<const int> owners = map.owners();
    std::vector<std::int32_t> perm(owners.size());
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](std::int32_t a, std::int32_t b) { return owners[a] < owners[b]; });
    _displs_remote.assign(_src.size() + 1, 0);
    _sizes_remote.assign(_src.size(), 0);
    Comm::Buffers send(_src.size());
    for (std::int32_t k : perm)
    {
      const std::size_t i = std::ranges::lower_bound(_src, owners[k]) - _src.begin();
      send[i].push_back(map.ghosts()[k]);
    }
    for (std::size_t i = 0; i < _src.size(); ++i)
    {
      _sizes_remote[i] = (int)send[i].size();
      _displs_remote[i + 1] = _displs_remote[i] + _sizes_remote[i];
    }
    // ghost global indices to their owners (:142-159)
    const Comm::Buffers recv = map.comm().neighbor_alltoallv(_src, _dest, send);
    _displs_local.assign(_dest.size() + 1, 0);
    _sizes_local.assign(_dest.size(), 0);
    const std::array<std::int64_t, 2> range = map.local_range();
    for (std::size_t j = 0; j < _dest.size(); ++j)
    {
      _sizes_local[j] = (int)recv[j].size();
      _displs_local[j + 1] = _displs_local[j] + _sizes_local[j];
      for (std::int64_t g : recv[j])
      {
        if (g < range[0] or g >= range[1])
          throw std::runtime_error("Scatterer: received index outside the owned range");
        for (int c = 0; c < bs; ++c) // expanded by the block size (:171-197)
          _local_inds.push_back((std::int32_t)((g - range[0]) * bs + c));
      }
    }
    for (std::int32_t k : perm)
      for (int c = 0; c < bs; ++c)
        _remote_inds.push_back(k * bs + c);
    for (auto* v : {&_sizes_local, &_displs_local, &_sizes_remote, &_displs_remote})
      for (int& e : *v)
        e *= bs;
  }
  /// Several ranks: the arrays computed by the reference constructor (common/Scatterer.h:84-197)
  Scatterer(const IndexMap& map, std::vector<std::int32_t> local_inds, std::vector<std::int32_t> remote_inds,
            std::vector<int> sizes_local, std::vector<int> displs_local, std::vector<int> sizes_remote,
            std::vector<int> displs_remote)
      : _src(map.src().begin(), map.src().end()), _dest(map.dest().begin(), map.dest().end()),
        _remote_inds(std::move(remote_inds)), _sizes_remote(std::move(sizes_remote)),
        _displs_remote(std::move(displs_remote)), _local_inds(std::move(local_inds)),
        _sizes_local(std::move(sizes_local)), _displs_local(std::move(displs_local)), _nccl(map.comm().nccl)
  {
  }
  Scatterer(const Scatterer&) = delete;
  ~Scatterer() { bfx_scatter_destroy(_plan); }
  const std::vector<std::int32_t>& local_indices() const noexcept { return _local_inds; }
  const std::vector<std::int32_t>& remote_indices() const noexcept { return _remote_inds; }
  const std::vector<int>& sizes_local() const noexcept { return _sizes_local; }
  const std::vector<int>& displs_local() const noexcept { return _displs_local; }
  const std::vector<int>& sizes_remote() const noexcept { return _sizes_remote; }
  const std::vector<int>& displs_remote() const noexcept { return _displs_remote; }
  std::size_t num_p2p_requests() const noexcept { return _dest.size() + _src.size(); }
  /// the device side of the plan (index arrays, staging buffers, stream, events), created by the first exchange
  bfx_scatter_t* plan() const
  {
    if (!_plan)
      const_cast<Scatterer*>(this)->create(_nccl);
    return _plan;
  }

private:
  void create(bfx_comm_t* nccl)
  {
    if (_displs_local.empty())
      _displs_local.assign(1, 0);
    if (_displs_remote.empty())
      _displs_remote.assign(1, 0);
    check(bfx_scatter_create(&_plan, nccl, _local_inds.data(), (std::int64_t)_local_inds.size(), _remote_inds.data(),
                             (std::int64_t)_remote_inds.size(), _sizes_local.data(), _displs_local.data(), _dest.data(),
                             (int)_dest.size(), _sizes_remote.data(), _displs_remote.data(), _src.data(),
                             (int)_src.size()));
  }
  std::vector<int> _src, _dest;
  std::vector<std::int32_t> _remote_inds;
  std::vector<int> _sizes_remote, _displs_remote;
  std::vector<std::int32_t> _local_inds;
  std::vector<int> _sizes_local, _displs_local;
  bfx_comm_t* _nccl = nullptr;
  bfx_scatter_t* _plan = nullptr;
};
} // namespace common

// -------------------------------------------------------------------------------------------------
namespace la
{
enum class Norm
{
  l1,
  l2,
  linf,
  frobenius
};

/// la::SparsityPattern (la/SparsityPattern.cpp:194-240, 264-491)
class SparsityPattern
{
public:
  SparsityPattern(common::Comm comm, std::array<std::shared_ptr<const common::IndexMap>, 2> maps, std::array<int, 2> bs)
      : _comm(comm), _index_maps(std::move(maps)), _bs(bs)
  {
  }
  void insert(std::int32_t row, std::int32_t col)
  {
    closed();
    _cache_rows.push_back(row);
    _cache_cols.push_back(col);
  }
  void insert(std::span<const std::int32_t> rows, std::span<const std::int32_t> cols)
  {
    closed();
    for (std::int32_t row : rows)
    {
      _cache_rows.insert(_cache_rows.end(), cols.size(), row);
      _cache_cols.insert(_cache_cols.end(), cols.begin(), cols.end());
    }
  }
  void insert_diagonal(std::span<const std::int32_t> rows)
  {
    closed();
    _cache_rows.insert(_cache_rows.end(), rows.begin(), rows.end());
    _cache_cols.insert(_cache_cols.end(), rows.begin(), rows.end());
  }
  /// fem::sparsitybuild::cells with device dofmaps: deferred to the GPU builder (bfx_sparsity_build)
  void insert_cells_device(const std::int32_t* dofmap0_dev, int nd0, const std::int32_t* dofmap1_dev, int nd1,
                           std::int64_t ncells)
  {
    closed();
    _dm0 = dofmap0_dev, _dm1 = dofmap1_dev, _nd0 = nd0, _nd1 = nd1, _ncells = ncells;
  }
  /// SparsityPattern::finalize (la/SparsityPattern.cpp:264-491).  On several ranks (collective): the entries inserted
  /// into ghost rows travel to the row owners as (global row, global column, column owner) triplets (:291-383),
  /// columns the owner did not know are appended to its column map in arrival order (:389-423), then every row is
  /// de-duplicated and sorted (:438-478) and the column IndexMap is rebuilt (:488-490).
  void finalize()
  {
    if (_finalized)
      throw std::runtime_error("Sparsity pattern has already been finalised.");
    const common::IndexMap& m0 = *_index_maps[0];
    const common::IndexMap& m1 = *_index_maps[1];
    const std::int32_t local_size0 = m0.size_local(), n0 = local_size0 + m0.num_ghosts();
    const std::int32_t l1 = m1.size_local();
    std::vector<std::int64_t> col_ghosts(m1.ghosts().begin(), m1.ghosts().end());
    std::vector<int> col_owners(m1.owners().begin(), m1.owners().end());
    if (_comm.size > 1)
    {
      if (_dm0)
        throw std::runtime_error("SparsityPattern::finalize on several ranks: insert the cells on the host "
                                 "(the device cell list is the single-rank fast path of this mirror)");
      if (!_comm.neighbor_alltoallv)
        throw std::runtime_error("SparsityPattern::finalize on several ranks needs Comm::neighbor_alltoallv");
      // distinct entries of every ghost row, columns in the order of their first insertion (the COO cache order)
      std::vector<std::vector<std::int32_t>> grow(m0.num_ghosts());
      for (std::size_t k = 0; k < _cache_rows.size(); ++k)
      {
        const std::int32_t r = _cache_rows[k];
        if (r >= local_size0 and r < n0)
        {
          auto& v = grow[r - local_size0];
          if (std::find(v.begin(), v.end(), _cache_cols[k]) == v.end())
            v.push_back(_cache_cols[k]);
        }
      }
      const std::span<const int> src = m0.src(), dest = m0.dest();
      common::Comm::Buffers send(src.size());
      const std::array<std::int64_t, 2> range1 = m1.local_range();
      for (std::int32_t g = 0; g < m0.num_ghosts(); ++g) // per owner: ghost rows ascending
      {
        auto& out = send[std::ranges::lower_bound(src, m0.owners()[g]) - src.begin()];
        for (std::int32_t c : grow[g])
        {
          const bool owned = c < l1;
          out.push_back(m0.ghosts()[g]);
          out.push_back(owned ? range1[0] + c : m1.ghosts()[c - l1]);
          out.push_back(owned ? _comm.rank : m1.owners()[c - l1]);
        }
      }
      const common::Comm::Buffers recv = _comm.neighbor_alltoallv(src, dest, send);
      // received columns -> local indices; unknown ones join the ghost list in arrival order
      std::unordered_map<std::int64_t, std::int32_t> ghost_local;
      for (std::size_t i = 0; i < col_ghosts.size(); ++i)
        ghost_local.emplace(col_ghosts[i], l1 + (std::int32_t)i);
      const std::int64_t row0 = m0.local_range()[0];
      for (const auto& buf : recv)
        for (std::size_t k = 0; k + 2 < buf.size(); k += 3)
        {
          const std::int64_t gr = buf[k], gc = buf[k + 1];
          std::int32_t lc;
          if (gc >= range1[0] and gc < range1[1])
            lc = (std::int32_t)(gc - range1[0]);
          else
          {
            auto [it, fresh] = ghost_local.try_emplace(gc, l1 + (std::int32_t)col_ghosts.size());
            if (fresh)
            {
              col_ghosts.push_back(gc);
              col_owners.push_back((int)buf[k + 2]);
            }
            lc = it->second;
          }
          _cache_rows.push_back((std::int32_t)(gr - row0));
          _cache_cols.push_back(lc);
        }
    }
    if (_dm0)
    {
      check(bfx_sparsity_build(&_csr, n0, local_size0, l1, _dm0, _nd0, _dm1, _nd1, nullptr, _ncells, _cache_rows.data(),
                               _cache_cols.data(), (std::int64_t)_cache_rows.size(), _bs[0], _bs[1], nullptr));
      const std::int64_t nnz = bfx_csr_nnz(_csr);
      _offsets.resize(n0 + 1);
      _edges.resize(nnz);
      std::vector<std::int64_t> od(n0);
      check(bfx_csr_get_structure(_csr, _offsets.data(), _edges.data(), od.data()));
      _off_diagonal_offsets.resize(n0);
      for (std::int32_t i = 0; i < n0; ++i)
        _off_diagonal_offsets[i] = (std::int32_t)(od[i] - _offsets[i]);
    }
    else
    {
      // bucket by row, de-duplicate, sort (SparsityPattern.cpp:438-478); the device copy is made on first use
      std::vector<std::vector<std::int32_t>> rows(n0);
      for (std::size_t k = 0; k < _cache_rows.size(); ++k)
      {
        if (_cache_rows[k] < 0 or _cache_rows[k] >= n0)
          throw std::runtime_error("SparsityPattern: row out of range");
        rows[_cache_rows[k]].push_back(_cache_cols[k]);
      }
      _offsets.assign(1, 0);
      _off_diagonal_offsets.clear();
      for (auto& r : rows)
      {
        std::ranges::sort(r);
        r.erase(std::unique(r.begin(), r.end()), r.end());
        _off_diagonal_offsets.push_back((std::int32_t)std::distance(r.begin(), std::ranges::lower_bound(r, l1)));
        _edges.insert(_edges.end(), r.begin(), r.end());
        _offsets.push_back(_offsets.back() + (std::int64_t)r.size());
      }
    }
    std::vector<std::int32_t>().swap(_cache_rows);
    std::vector<std::int32_t>().swap(_cache_cols);
    if (_comm.size > 1) // the column map with the ghosts found (collective, like the reference's IndexMap constructor)
      _index_maps[1] = std::make_shared<const common::IndexMap>(_comm, l1, col_ghosts, col_owners);
    _finalized = true;
  }
  std::shared_ptr<const common::IndexMap> index_map(int dim) const { return _index_maps.at(dim); }
  int block_size(int dim) const { return _bs[dim]; }
  std::int64_t num_nonzeros() const
  {
    final();
    return (std::int64_t)_edges.size();
  }
  std::pair<std::span<const std::int32_t>, std::span<const std::int64_t>> graph() const
  {
    final();
    return {_edges, _offsets};
  }
  std::span<const std::int32_t> off_diagonal_offsets() const
  {
    final();
    return _off_diagonal_offsets;
  }
  /// device structure shared with MatrixCSR (ownership passes to the first matrix built from it); patterns built on
  /// the host upload theirs here, on first use
  bfx_csr_t* release_csr() const
  {
    final();
    if (_csr_released)
      return nullptr;
    if (!_csr)
    {
      const std::int32_t n0 = _index_maps[0]->size_local() + _index_maps[0]->num_ghosts();
      std::vector<std::int64_t> od(n0);
      for (std::int32_t i = 0; i < n0; ++i)
        od[i] = _offsets[i] + _off_diagonal_offsets[i];
      check(bfx_csr_create(&_csr, n0, _index_maps[0]->size_local(), _offsets.data(), _edges.data(), od.data(), _bs[0], _bs[1]));
    }
    _csr_released = true;
    return _csr;
  }
  ~SparsityPattern()
  {
    if (!_csr_released)
      bfx_csr_destroy(_csr);
  }

private:
  void closed() const
  {
    if (_finalized)
      throw std::runtime_error("Cannot insert into sparsity pattern. It has already been finalized");
  }
  void final() const
  {
    if (!_finalized)
      throw std::runtime_error("Sparsity pattern has not been finalised.");
  }
  common::Comm _comm;
  std::array<std::shared_ptr<const common::IndexMap>, 2> _index_maps;
  std::array<int, 2> _bs;
  std::vector<std::int32_t> _cache_rows, _cache_cols;
  const std::int32_t *_dm0 = nullptr, *_dm1 = nullptr;
  int _nd0 = 0, _nd1 = 0;
  std::int64_t _ncells = 0;
  std::vector<std::int32_t> _edges, _off_diagonal_offsets;
  std::vector<std::int64_t> _offsets;
  mutable bfx_csr_t* _csr = nullptr;
  mutable bool _csr_released = false;
  bool _finalized = false;
};

/// la::Vector<double> with device storage
template <typename T = double>
class Vector
{
  static_assert(std::is_same_v<T, double>, "the B200 path computes in fp64");

public:
  using value_type = T;
  Vector(std::shared_ptr<const common::IndexMap> map, int bs)
      : _map(map), _bs(bs), _x((std::size_t)bs * (map->size_local() + map->num_ghosts())),
        _scatterer(std::make_shared<common::Scatterer>(*map, bs))
  {
  }
  Vector(std::shared_ptr<const common::IndexMap> map, int bs, std::shared_ptr<const common::Scatterer> sc)
      : _map(map), _bs(bs), _x((std::size_t)bs * (map->size_local() + map->num_ghosts())), _scatterer(std::move(sc))
  {
  }
  void set(T v) { check(bfx_fill((std::int64_t)_x.size(), v, _x.data(), nullptr)); }
  void scatter_fwd_begin() { check(bfx_scatter_fwd_begin(_scatterer->plan(), _x.data(), nullptr)); }
  void scatter_fwd_end()
  {
    check(bfx_scatter_fwd_end(_scatterer->plan(), _x.data(), (std::int64_t)_bs * _map->size_local(), nullptr));
  }
  void scatter_fwd()
  {
    scatter_fwd_begin();
    scatter_fwd_end();
  }
  void scatter_rev_begin()
  {
    check(bfx_scatter_rev_begin(_scatterer->plan(), _x.data(), (std::int64_t)_bs * _map->size_local(), nullptr));
  }
  /// op: std::plus<T> => add; anything else => insert (la/Vector.h:96-114)
  template <class BinaryOperation>
  void scatter_rev(BinaryOperation)
  {
    scatter_rev_begin();
    constexpr int op = std::is_same_v<BinaryOperation, std::plus<T>> ? 1 : 0;
    check(bfx_scatter_rev_end(_scatterer->plan(), _x.data(), op, nullptr));
  }
  std::shared_ptr<const common::IndexMap> index_map() const { return _map; }
  constexpr int bs() const { return _bs; }
  DeviceArray<T>& array() { return _x; }
  const DeviceArray<T>& array() const { return _x; }

private:
  std::shared_ptr<const common::IndexMap> _map;
  int _bs;
  DeviceArray<T> _x;
  std::shared_ptr<const common::Scatterer> _scatterer;
};

/// la::inner_product (la/Vector.h:434-460); the all-reduce over ranks runs on the device communicator
template <class V>
auto inner_product(const V& a, const V& b)
{
  const std::int64_t n = (std::int64_t)a.bs() * a.index_map()->size_local();
  if (n != (std::int64_t)b.bs() * b.index_map()->size_local())
    throw std::runtime_error("Incompatible vector sizes");
  double local = 0;
  check(bfx_dot(n, a.array().data(), b.array().data(), &local, nullptr));
  if (a.index_map()->comm().size > 1)
  {
    DeviceArray<double> d(1);
    d.copy_from(std::span<const double>(&local, 1));
    check(bfx_comm_allreduce(a.index_map()->comm().nccl, d.data(), 1, 0, nullptr));
    local = d.to_host()[0];
  }
  return local;
}
template <class V>
auto squared_norm(const V& a)
{
  return inner_product(a, a);
}
/// la::norm (la/Vector.h:479-514)
template <class V>
auto norm(const V& x, Norm type = Norm::l2)
{
  const std::int64_t n = (std::int64_t)x.bs() * x.index_map()->size_local();
  double r = 0;
  switch (type)
  {
  case Norm::l2: return std::sqrt(squared_norm(x));
  case Norm::l1: check(bfx_norm(n, x.array().data(), 0, &r, nullptr)); break;
  case Norm::linf: check(bfx_norm(n, x.array().data(), 2, &r, nullptr)); break;
  default: throw std::runtime_error("Norm type not supported");
  }
  if (x.index_map()->comm().size > 1)
  {
    DeviceArray<double> d(1);
    d.copy_from(std::span<const double>(&r, 1));
    check(bfx_comm_allreduce(x.index_map()->comm().nccl, d.data(), 1, type == Norm::linf ? 1 : 0, nullptr));
    r = d.to_host()[0];
  }
  return r;
}

namespace impl
{
/// la::impl::Sparsity (la/matmul.h): a finalised pattern given by its arrays - what la::transpose / la::matmul hand
/// to the MatrixCSR constructor (la/mattrans.h:153-155, 431-434)
struct Sparsity
{
  std::shared_ptr<const common::IndexMap> row_map, col_map;
  std::vector<std::int32_t> cols;
  std::vector<std::int64_t> offsets;
  std::vector<std::int32_t> off_diag;
  std::array<int, 2> bs;
  std::shared_ptr<const common::IndexMap> index_map(int dim) const { return dim == 0 ? row_map : col_map; }
  int block_size(int dim) const { return bs.at(dim); }
  std::pair<std::span<const std::int32_t>, std::span<const std::int64_t>> graph() const { return {cols, offsets}; }
  std::span<const std::int32_t> off_diagonal_offsets() const { return off_diag; }
  bfx_csr_t* release_csr() const { return nullptr; }
};
} // namespace impl

/// la::MatrixCSR<double> (compact block mode) with device storage
namespace impl
{
/// The ghost-row exchange plan the MatrixCSR constructor builds (la/MatrixCSR.h:705-849), members as in the reference:
/// ghost_row_to_rank (index into src), val_send_disp / val_recv_disp (already x bs0*bs1), unpack_pos (block positions
/// of the received entries in the owner's CSR).
struct GhostRowPlan
{
  std::vector<std::int32_t> ghost_row_to_rank;
  std::vector<std::int64_t> val_send_disp, val_recv_disp, unpack_pos;
};

/// Collective over the neighbourhood of the row map (one exchange of (global row, global column) pairs)
inline GhostRowPlan matrix_ghost_plan(const common::IndexMap& m0, const common::IndexMap& m1, std::array<int, 2> bs,
                                      std::span<const std::int64_t> row_ptr, std::span<const std::int32_t> cols)
{
  const common::Comm& comm = m0.comm();
  if (!comm.neighbor_alltoallv)
    throw std::runtime_error("MatrixCSR on several ranks needs Comm::neighbor_alltoallv");
  const std::int64_t bs2 = (std::int64_t)bs[0] * bs[1];
  const std::int32_t ls0 = m0.size_local(), ls1 = m1.size_local();
  const std::span<const int> src = m0.src(), dest = m0.dest();
  GhostRowPlan p;
  // owner (as index into src) of every ghost row; entries per owner (:725-742)
  std::vector<std::int64_t> per_proc(src.size(), 0);
  for (std::int32_t i = 0; i < m0.num_ghosts(); ++i)
  {
    const std::int32_t r = (std::int32_t)(std::ranges::lower_bound(src, m0.owners()[i]) - src.begin());
    p.ghost_row_to_rank.push_back(r);
    per_proc[r] += row_ptr[ls0 + i + 1] - row_ptr[ls0 + i];
  }
  p.val_send_disp.assign(src.size() + 1, 0);
  std::partial_sum(per_proc.begin(), per_proc.end(), p.val_send_disp.begin() + 1);
  // (global row, global column) of every ghost-row entry, per owner in ghost-row order (:750-775)
  common::Comm::Buffers send(src.size());
  for (std::int32_t i = 0; i < m0.num_ghosts(); ++i)
  {
    auto& out = send[p.ghost_row_to_rank[i]];
    for (std::int64_t j = row_ptr[ls0 + i]; j < row_ptr[ls0 + i + 1]; ++j)
    {
      out.push_back(m0.ghosts()[i]);
      const std::int32_t c = cols[j];
      out.push_back(c < ls1 ? m1.local_range()[0] + c : m1.ghosts()[c - ls1]);
    }
  }
  const common::Comm::Buffers recv = comm.neighbor_alltoallv(src, dest, send); // :777-800
  // positions in the owner's CSR (:813-846)
  std::unordered_map<std::int64_t, std::int32_t> ghost_local;
  for (std::int32_t i = 0; i < m1.num_ghosts(); ++i)
    ghost_local.emplace(m1.ghosts()[i], ls1 + i);
  p.val_recv_disp.assign(1, 0);
  for (const auto& buf : recv)
  {
    for (std::size_t k = 0; k + 1 < buf.size(); k += 2)
    {
      const std::int64_t lrow = buf[k] - m0.local_range()[0];
      if (lrow < 0 or lrow >= ls0)
        throw std::runtime_error("MatrixCSR: received a ghost row this rank does not own");
      std::int32_t lcol;
      if (buf[k + 1] >= m1.local_range()[0] and buf[k + 1] < m1.local_range()[1])
        lcol = (std::int32_t)(buf[k + 1] - m1.local_range()[0]);
      else if (auto it = ghost_local.find(buf[k + 1]); it != ghost_local.end())
        lcol = it->second;
      else
        throw std::runtime_error("MatrixCSR: received ghost-row entry not in sparsity");
      const auto b = cols.begin() + row_ptr[lrow], e = cols.begin() + row_ptr[lrow + 1];
      const auto it = std::lower_bound(b, e, lcol);
      if (it == e or *it != lcol)
        throw std::runtime_error("MatrixCSR: received ghost-row entry not in sparsity");
      p.unpack_pos.push_back(std::distance(cols.begin(), it));
    }
    p.val_recv_disp.push_back(p.val_recv_disp.back() + bs2 * (std::int64_t)(buf.size() / 2));
  }
  for (auto& d : p.val_send_disp)
    d *= bs2;
  return p;
}
} // namespace impl

template <typename T = double>
class MatrixCSR
{
  static_assert(std::is_same_v<T, double>, "the B200 path computes in fp64");

public:
  using value_type = T;
  /// MatrixCSR(const SparsityPattern&) — la/MatrixCSR.h:628-703 (+ single-rank ghost plan: empty)
  explicit MatrixCSR(const SparsityPattern& p) { init(p); }
  /// MatrixCSR(const impl::Sparsity&) with its values (host), as la::transpose builds its result (la/mattrans.h:155-157)
  MatrixCSR(const impl::Sparsity& p, std::span<const T> values)
  {
    init(p);
    _data.copy_from(values);
  }
  MatrixCSR(const MatrixCSR&) = delete;
  ~MatrixCSR()
  {
    bfx_csr_scatter_destroy(_scatter);
    bfx_csr_destroy(_csr);
  }

private:
  template <class Pattern>
  void init(const Pattern& p)
  {
    _index_maps = {p.index_map(0), p.index_map(1)};
    _bs = {p.block_size(0), p.block_size(1)};
    _cols.assign(p.graph().first.begin(), p.graph().first.end());
    _row_ptr.assign(p.graph().second.begin(), p.graph().second.end());
    std::span<const std::int32_t> nd = p.off_diagonal_offsets();
    _off_diagonal_offset.resize(nd.size());
    for (std::size_t i = 0; i < nd.size(); ++i)
      _off_diagonal_offset[i] = _row_ptr[i] + nd[i];
    _csr = p.release_csr();
    if (!_csr)
    {
      const std::int32_t n0 = _index_maps[0]->size_local() + _index_maps[0]->num_ghosts();
      check(bfx_csr_create(&_csr, n0, _index_maps[0]->size_local(), _row_ptr.data(), _cols.data(),
                           _off_diagonal_offset.data(), _bs[0], _bs[1]));
    }
    _data = DeviceArray<T>(_cols.size() * _bs[0] * _bs[1]);
    if (_index_maps[0]->comm().size > 1) // the constructor builds the ghost-row plan (collective, la/MatrixCSR.h:705-849)
    {
      const impl::GhostRowPlan g = impl::matrix_ghost_plan(*_index_maps[0], *_index_maps[1], _bs, _row_ptr, _cols);
      set_ghost_plan(g.ghost_row_to_rank, g.val_send_disp, g.val_recv_disp, g.unpack_pos);
    }
  }

public:
  /// Install the ghost-row exchange plan computed by the reference constructor (la/MatrixCSR.h:705-849)
  void set_ghost_plan(std::span<const std::int32_t> ghost_row_to_rank, std::span<const std::int64_t> val_send_disp,
                      std::span<const std::int64_t> val_recv_disp, std::span<const std::int64_t> unpack_pos)
  {
    std::vector<std::int32_t> src(_index_maps[0]->src().begin(), _index_maps[0]->src().end());
    std::vector<std::int32_t> dest(_index_maps[0]->dest().begin(), _index_maps[0]->dest().end());
    check(bfx_csr_scatter_create(&_scatter, _csr, _index_maps[0]->comm().nccl, ghost_row_to_rank.data(),
                                 (std::int32_t)ghost_row_to_rank.size(), val_send_disp.data(), src.data(), (int)src.size(),
                                 val_recv_disp.data(), dest.data(), (int)dest.size(), unpack_pos.data()));
  }
  /// MatrixCSR::set(value) (la/MatrixCSR.h:239-241); a zeroed matrix lets the next assembly overwrite instead of add
  void set(T x)
  {
    check(bfx_fill((std::int64_t)_data.size(), x, _data.data(), nullptr));
    _known_zero = x == T(0);
  }
  /// true until something is added / set after set(0) (or construction): the assembler's BFX_VALUES_OVERWRITE switch
  bool known_zero() const { return _known_zero; }
  void touched() { _known_zero = false; }
  /// The insertion functor the reference passes to fem::assemble_matrix (la/MatrixCSR.h:134-175 mat_add_values):
  /// here a handle on the matrix, consumed by the fem::assemble_matrix overload below (the element tensors never
  /// visit the host, so there is nothing to call back per cell)
  struct MatAdd
  {
    MatrixCSR* A;
  };
  MatAdd mat_add_values() { return MatAdd{this}; }
  /// MatrixCSR::set / add<BS0,BS1> — la/MatrixCSR.h:265-335
  template <int BS0 = 1, int BS1 = 1>
  void set(std::span<const T> x, std::span<const std::int32_t> rows, std::span<const std::int32_t> cols)
  {
    insert<BS0, BS1>(x, rows, cols, 0);
  }
  template <int BS0 = 1, int BS1 = 1>
  void add(std::span<const T> x, std::span<const std::int32_t> rows, std::span<const std::int32_t> cols)
  {
    insert<BS0, BS1>(x, rows, cols, 1);
  }
  std::int32_t num_owned_rows() const { return _index_maps[0]->size_local(); }
  std::int32_t num_all_rows() const { return (std::int32_t)_row_ptr.size() - 1; }
  void scatter_rev_begin()
  {
    if (_scatter)
      check(bfx_csr_scatter_rev_begin(_scatter, _data.data(), nullptr));
  }
  void scatter_rev_end()
  {
    if (_scatter)
      check(bfx_csr_scatter_rev_end(_scatter, _data.data(), nullptr));
  }
  void scatter_rev()
  {
    scatter_rev_begin();
    scatter_rev_end();
  }
  double squared_norm() const
  {
    double r = 0;
    check(bfx_csr_squared_norm(_csr, _data.data(), &r, nullptr));
    if (const common::Comm& c = _index_maps[0]->comm(); c.size > 1) // la/MatrixCSR.h:483-484
    {
      DeviceArray<double> d(1);
      d.copy_from(std::span<const double>(&r, 1));
      check(bfx_comm_allreduce(c.nccl, d.data(), 1, 0, nullptr));
      r = d.to_host()[0];
    }
    return r;
  }
  /// y += A x — la/MatrixCSR.h:877-946 (split around the ghost update of x)
  void mult(Vector<T>& x, Vector<T>& y) const
  {
    if (_index_maps[0]->comm().size == 1)
    {
      check(bfx_spmv(_csr, _data.data(), x.array().data(), y.array().data(), BFX_SPMV_FULL, nullptr));
      return;
    }
    x.scatter_fwd_begin();
    check(bfx_spmv(_csr, _data.data(), x.array().data(), y.array().data(), BFX_SPMV_DIAG, nullptr));
    x.scatter_fwd_end();
    check(bfx_spmv(_csr, _data.data(), x.array().data(), y.array().data(), BFX_SPMV_OFFDIAG, nullptr));
  }
  /// y += A^T x — la/MatrixCSR.h:950-1016
  void multT(Vector<T>& x, Vector<T>& y) const
  {
    check(bfx_spmvT(_csr, _data.data(), x.array().data(), y.array().data(), BFX_SPMV_OFFDIAG, nullptr));
    y.scatter_rev(std::plus<T>{});
    check(bfx_spmvT(_csr, _data.data(), x.array().data(), y.array().data(), BFX_SPMV_DIAG, nullptr));
  }
  std::vector<T> to_dense() const
  {
    const std::size_t nrows = num_all_rows();
    const std::size_t ncols = _index_maps[1]->size_local() + _index_maps[1]->num_ghosts();
    std::vector<T> A(nrows * ncols * _bs[0] * _bs[1], 0), v = _data.to_host();
    for (std::size_t r = 0; r < nrows; ++r)
      for (std::int64_t j = _row_ptr[r]; j < _row_ptr[r + 1]; ++j)
        for (int i0 = 0; i0 < _bs[0]; ++i0)
          for (int i1 = 0; i1 < _bs[1]; ++i1)
            A[(r * _bs[0] + i0) * ncols * _bs[1] + _cols[j] * _bs[1] + i1] = v[j * _bs[0] * _bs[1] + i0 * _bs[1] + i1];
    return A;
  }
  std::shared_ptr<const common::IndexMap> index_map(int dim) const { return _index_maps.at(dim); }
  /// mutable access: the caller may write, so the matrix no longer counts as known-zero
  DeviceArray<T>& values()
  {
    _known_zero = false;
    return _data;
  }
  const DeviceArray<T>& values() const { return _data; }
  const std::vector<std::int64_t>& row_ptr() const { return _row_ptr; }
  const std::vector<std::int32_t>& cols() const { return _cols; }
  const std::vector<std::int64_t>& off_diag_offset() const { return _off_diagonal_offset; }
  std::array<int, 2> block_size() const { return _bs; }
  const bfx_csr_t* csr() const { return _csr; }

private:
  template <int BS0, int BS1>
  void insert(std::span<const T> x, std::span<const std::int32_t> rows, std::span<const std::int32_t> cols, int op)
  {
    int kind;
    if (_bs[0] == BS0 and _bs[1] == BS1)
      kind = 0;
    else if (_bs[0] == 1 and _bs[1] == 1)
      kind = 1;
    else if (BS0 == 1 and BS1 == 1)
      kind = 2;
    else
      throw std::runtime_error("Unsupported block size in MatrixCSR insertion");
    check(bfx_csr_insert(_csr, _data.data(), kind, BS0, BS1, x.data(), rows.data(), (int)rows.size(), cols.data(),
                         (int)cols.size(), op, nullptr));
    _known_zero = false;
  }
  std::array<std::shared_ptr<const common::IndexMap>, 2> _index_maps;
  std::array<int, 2> _bs;
  DeviceArray<T> _data;
  std::vector<std::int32_t> _cols;
  std::vector<std::int64_t> _row_ptr, _off_diagonal_offset;
  bfx_csr_t* _csr = nullptr;
  bfx_csr_scatter_t* _scatter = nullptr;
  bool _known_zero = true; // fresh values are zero (DeviceArray)
};

/// la::transpose (la/mattrans.h:121-159): the branch without neighbours in the column map (one rank, or no ghost
/// columns anywhere near this rank); impl::local_transpose runs on the device, bit-exact (bfx_csr_transpose_local).
/// With ghost columns the entries travel to the column owners first - host-side integer work the caller keeps
/// (dolfinx_b200.la.matrix_transpose_plan is the tested restatement of la/mattrans.h:200-434).
template <typename T>
MatrixCSR<T> transpose(const MatrixCSR<T>& A)
{
  auto m0 = A.index_map(0);
  auto m1 = A.index_map(1);
  if (m0->comm().size != 1 and !(m1->src().empty() and m1->dest().empty()))
    throw std::runtime_error("la::transpose: ghost columns present - exchange them first (la/mattrans.h:200-434)");
  const std::array<int, 2> bs = A.block_size();
  const std::int32_t n_row = m0->size_local(), n_col = m1->size_local();
  const std::int64_t cap = A.row_ptr()[n_row];
  DeviceArray<std::int64_t> rp(n_col + 1);
  DeviceArray<std::int32_t> cols(std::max<std::int64_t>(cap, 1));
  DeviceArray<T> vals(std::max<std::int64_t>(cap, 1) * bs[0] * bs[1]);
  std::int64_t nnz = 0;
  check(bfx_csr_transpose_local(A.csr(), A.values().data(), n_col, rp.data(), cols.data(), vals.data(), cap, &nnz, nullptr));
  impl::Sparsity sp;
  sp.row_map = std::make_shared<common::IndexMap>(m0->comm(), n_col); // la/mattrans.h:142-143
  sp.col_map = std::make_shared<common::IndexMap>(m0->comm(), n_row);
  sp.offsets = rp.to_host();
  sp.cols = cols.to_host();
  sp.cols.resize(nnz);
  sp.off_diag.resize(n_col);
  for (std::int32_t j = 0; j < n_col; ++j)
    sp.off_diag[j] = static_cast<std::int32_t>(sp.offsets[j + 1] - sp.offsets[j]); // all columns owned (:148-151)
  sp.bs = {bs[1], bs[0]};
  std::vector<T> v = vals.to_host();
  return MatrixCSR<T>(sp, std::span<const T>(v.data(), nnz * bs[0] * bs[1]));
}
} // namespace la

// -------------------------------------------------------------------------------------------------
namespace fem
{
enum class IntegralType : std::int8_t
{
  cell = 0,
  exterior_facet = 1,
  interior_facet = 2,
  vertex = 3
};

/// mesh::Geometry::x() / dofmap() — the two arrays the assembler reads (mesh/Geometry.h:131,155)
struct Mesh
{
  Mesh(common::Comm comm, std::span<const double> x, std::span<const std::int32_t> x_dofmap, int nx)
      : comm(comm), nx(nx), num_cells((std::int64_t)x_dofmap.size() / nx), num_nodes((std::int64_t)x.size() / 3), x(x),
        x_dofmap(x_dofmap)
  {
  }
  common::Comm comm;
  int nx;
  std::int64_t num_cells, num_nodes;
  DeviceArray<double> x;
  DeviceArray<std::int32_t> x_dofmap;
};

/// fem::DofMap accessor subset (fem/DofMap.h:127-167)
struct DofMap
{
  DofMap(std::span<const std::int32_t> map, int nd, int bs, std::shared_ptr<const common::IndexMap> index_map)
      : nd(nd), _bs(bs), index_map(std::move(index_map)), _host(map.begin(), map.end()), dev(map)
  {
  }
  std::span<const std::int32_t> cell_dofs(std::int32_t c) const { return {_host.data() + (std::size_t)c * nd, (std::size_t)nd}; }
  int bs() const { return _bs; }
  int index_map_bs() const { return _bs; }
  int nd, _bs;
  std::shared_ptr<const common::IndexMap> index_map;
  std::vector<std::int32_t> _host;
  DeviceArray<std::int32_t> dev;
};

struct FunctionSpace
{
  FunctionSpace(std::shared_ptr<const Mesh> mesh, std::shared_ptr<const DofMap> dofmap)
      : _mesh(std::move(mesh)), _dofmap(std::move(dofmap))
  {
  }
  std::shared_ptr<const Mesh> mesh() const { return _mesh; }
  std::shared_ptr<const DofMap> dofmap() const { return _dofmap; }
  bool contains(const FunctionSpace& V) const { return this == &V; } // fem/FunctionSpace.h:153
  std::shared_ptr<const Mesh> _mesh;
  std::shared_ptr<const DofMap> _dofmap;
};

template <typename T = double>
struct Function
{
  explicit Function(std::shared_ptr<const FunctionSpace> V)
      : _V(V), _x(std::make_shared<la::Vector<T>>(V->dofmap()->index_map, V->dofmap()->bs()))
  {
  }
  std::shared_ptr<const FunctionSpace> function_space() const { return _V; }
  std::shared_ptr<la::Vector<T>> x() const { return _x; }
  std::shared_ptr<const FunctionSpace> _V;
  std::shared_ptr<la::Vector<T>> _x;
};

template <typename T = double>
struct Constant
{
  explicit Constant(T c) : value({c}) {}
  explicit Constant(std::span<const T> c) : value(c.begin(), c.end()) {}
  std::vector<T> value;
};

/// fem::integral_data (fem/Form.h:52-87) with a libbfx kernel id in place of the FFCx function pointer
struct integral_data
{
  int kernel;                          // BFX_K_*
  std::vector<std::int32_t> entities;  // cells, or flat (cell, local_facet) pairs; empty = all owned cells
  std::vector<int> coeffs;             // indices of the active coefficients
};

/// fem::Form built by hand (cf. cpp/demo/custom_kernel/main.cpp:77-81)
template <typename T = double>
class Form
{
public:
  Form(std::vector<std::shared_ptr<const FunctionSpace>> V,
       std::map<std::pair<IntegralType, int>, integral_data> integrals,
       std::vector<std::shared_ptr<const Function<T>>> coefficients = {},
       std::vector<std::shared_ptr<const Constant<T>>> constants = {})
      : _function_spaces(std::move(V)), _integrals(std::move(integrals)), _coefficients(std::move(coefficients)),
        _constants(std::move(constants))
  {
  }
  Form(const Form&) = delete; // fem/Form.h:341-353
  ~Form()
  {
    for (auto& [k, p] : _plans)
      bfx_asm_destroy(p);
  }
  int rank() const { return (int)_function_spaces.size(); }
  std::shared_ptr<const Mesh> mesh() const { return _function_spaces[0]->mesh(); }
  const std::vector<std::shared_ptr<const FunctionSpace>>& function_spaces() const { return _function_spaces; }
  const std::map<std::pair<IntegralType, int>, integral_data>& integrals() const { return _integrals; }
  const std::vector<std::shared_ptr<const Function<T>>>& coefficients() const { return _coefficients; }
  const std::vector<std::shared_ptr<const Constant<T>>>& constants() const { return _constants; }
  /// assembly plan of one integral, bound to a matrix structure (nullptr for linear forms / lifting)
  bfx_asm_t* plan(std::pair<IntegralType, int> key, const bfx_csr_t* csr) const
  {
    auto pk = std::make_pair(key, (const void*)csr);
    if (auto it = _plans.find(pk); it != _plans.end())
      return it->second;
    const integral_data& id = _integrals.at(key);
    auto m = mesh();
    auto dm0 = _function_spaces[0]->dofmap();
    auto dm1 = rank() == 2 ? _function_spaces[1]->dofmap() : nullptr;
    const bool cells = key.first == IntegralType::cell;
    const std::int64_t n = cells ? (id.entities.empty() ? m->num_cells : (std::int64_t)id.entities.size()) : 0;
    bfx_asm_t* p = nullptr;
    const std::int32_t nrows = dm0->index_map->size_local() + dm0->index_map->num_ghosts();
    check(bfx_asm_create(&p, csr, m->x_dofmap.data(), m->nx, dm0->dev.data(), dm0->nd, dm1 ? dm1->dev.data() : nullptr,
                         dm1 ? dm1->nd : 0, m->num_cells, (cells and !id.entities.empty()) ? id.entities.data() : nullptr,
                         n, nrows, 0, nullptr));
    _plans[pk] = p;
    return p;
  }
  /// scatter-add strategy of a cell integral of a bilinear form: the aggregated kernel of the element where it
  /// has one (built once per plan: chunk lists for the P1/P2 Poisson kernels, the transposed dofmap for Q1
  /// elasticity), else one fp64 RED per contribution
  int matrix_strategy(std::pair<IntegralType, int> key, const bfx_csr_t* csr) const
  {
    auto pk = std::make_pair(key, (const void*)csr);
    if (auto it = _strategies.find(pk); it != _strategies.end())
      return it->second;
    bfx_asm_t* p = plan(key, csr);
    const int kernel = _integrals.at(key).kernel;
    int strategy = BFX_ASM_ATOMIC;
    if (key.first == IntegralType::cell and rank() == 2)
    {
      const bool one_space = _function_spaces[0] == _function_spaces[1] and _function_spaces[0]->dofmap()->bs() == 1;
      const bool chunked = kernel == BFX_K_LAPLACE_P1_TRI_A or kernel == BFX_K_MASS_COEFF_P1_TRI_A
                           or kernel == BFX_K_POISSON_P1_TET_A or kernel == BFX_K_POISSON_P2_TET_A;
      int st = BFX_ERR_UNSUPPORTED;
      if (chunked)
      {
        st = bfx_asm_build_chunks(p, mesh()->x.data(), one_space ? BFX_CHUNKS_SYMMETRIC : 0, nullptr);
        if (st == BFX_OK)
          strategy = BFX_ASM_CHUNKED;
      }
      else if (kernel == BFX_K_ELASTICITY_Q1_HEX_A)
      {
        st = bfx_asm_build_rowgather(p, nullptr);
        if (st == BFX_OK)
          strategy = BFX_ASM_ROWGATHER;
      }
      if (st != BFX_OK and st != BFX_ERR_UNSUPPORTED)
        check(st);
    }
    _strategies[pk] = strategy;
    return strategy;
  }

private:
  std::vector<std::shared_ptr<const FunctionSpace>> _function_spaces;
  std::map<std::pair<IntegralType, int>, integral_data> _integrals;
  std::vector<std::shared_ptr<const Function<T>>> _coefficients;
  std::vector<std::shared_ptr<const Constant<T>>> _constants;
  mutable std::map<std::pair<std::pair<IntegralType, int>, const void*>, bfx_asm_t*> _plans;
  mutable std::map<std::pair<std::pair<IntegralType, int>, const void*>, int> _strategies;
};

/// fem::pack_constants (fem/pack.h:578-619)
template <typename T>
std::vector<T> pack_constants(const Form<T>& a)
{
  std::vector<T> c;
  for (auto& k : a.constants())
    c.insert(c.end(), k->value.begin(), k->value.end());
  return c;
}

/// fem::DirichletBC (fem/DirichletBC.h:262-601)
template <typename T = double>
class DirichletBC
{
public:
  /// dofs: sorted block indices; unrolled by the block size like the reference constructor (:357-361)
  DirichletBC(std::variant<std::shared_ptr<const Function<T>>, std::shared_ptr<const Constant<T>>> g,
              std::span<const std::int32_t> dofs, std::shared_ptr<const FunctionSpace> V)
      : _function_space(std::move(V)), _g(std::move(g))
  {
    const int bs = _function_space->dofmap()->bs();
    if (auto c = std::get_if<std::shared_ptr<const Constant<T>>>(&_g); c and (int)(*c)->value.size() != bs)
      throw std::runtime_error("Creating a DirichletBC using a Constant is not supported when the Constant size is "
                               "not equal to the block size of the constrained (sub-)space. Use a fem::Function to "
                               "create the fem::DirichletBC.");
    _dofs0.resize(dofs.size() * bs);
    for (std::size_t i = 0; i < dofs.size(); ++i)
      for (int k = 0; k < bs; ++k)
        _dofs0[bs * i + k] = bs * dofs[i] + k;
    const std::int32_t owned = bs * _function_space->dofmap()->index_map->size_local();
    _owned_indices0 = (std::int32_t)std::distance(_dofs0.begin(), std::ranges::lower_bound(_dofs0, owned));
    _dofs0_dev = DeviceArray<std::int32_t>(std::span<const std::int32_t>(_dofs0));
    if (auto c = std::get_if<std::shared_ptr<const Constant<T>>>(&_g))
      _g_dev = DeviceArray<T>(std::span<const T>((*c)->value));
  }
  std::shared_ptr<const FunctionSpace> function_space() const { return _function_space; }
  std::pair<std::span<const std::int32_t>, std::int32_t> dof_indices() const { return {_dofs0, _owned_indices0}; }
  const std::int32_t* dofs_dev() const { return _dofs0_dev.data(); }
  /// x[dof] = alpha (g[dof] - x0[dof]) — fem/DirichletBC.h:495-578; x / x0 are device arrays
  void set(DeviceArray<T>& x, const DeviceArray<T>* x0, T alpha = 1) const
  {
    const bool fn = std::holds_alternative<std::shared_ptr<const Function<T>>>(_g);
    const T* g = fn ? std::get<std::shared_ptr<const Function<T>>>(_g)->x()->array().data() : _g_dev.data();
    check(bfx_bc_set(x.data(), (std::int32_t)x.size(), _dofs0_dev.data(), nullptr, (std::int64_t)_dofs0.size(), g,
                     fn ? 0 : 1, _function_space->dofmap()->bs(), x0 ? x0->data() : nullptr, alpha, nullptr));
  }
  /// fem/DirichletBC.h:589-601; markers: device int8 array
  void mark_dofs(DeviceArray<std::int8_t>& markers) const
  {
    if (!_dofs0.empty() and *std::ranges::max_element(_dofs0) >= (std::int32_t)markers.size())
      throw std::runtime_error("Marker array is too short for the boundary condition dofs.");
    check(bfx_bc_mark(markers.data(), _dofs0_dev.data(), (std::int64_t)_dofs0.size(), nullptr));
  }

private:
  std::shared_ptr<const FunctionSpace> _function_space;
  std::variant<std::shared_ptr<const Function<T>>, std::shared_ptr<const Constant<T>>> _g;
  std::vector<std::int32_t> _dofs0;
  std::int32_t _owned_indices0 = 0;
  DeviceArray<std::int32_t> _dofs0_dev;
  DeviceArray<T> _g_dev;
};

/// fem::create_sparsity_pattern (fem/utils.h:197-218): all owned cells of the form's mesh
template <typename T>
la::SparsityPattern create_sparsity_pattern(const Form<T>& a)
{
  auto dm0 = a.function_spaces()[0]->dofmap();
  auto dm1 = a.function_spaces()[1]->dofmap();
  la::SparsityPattern sp(a.mesh()->comm, {dm0->index_map, dm1->index_map}, {dm0->index_map_bs(), dm1->index_map_bs()});
  sp.insert_cells_device(dm0->dev.data(), dm0->nd, dm1->dev.data(), dm1->nd, a.mesh()->num_cells);
  return sp;
}

namespace impl
{
template <typename T>
bfx_coeffs_t coeffs_for(const Form<T>& form, const integral_data& id)
{
  bfx_coeffs_t cf{};
  if (!id.coeffs.empty())
  {
    if (id.coeffs.size() != 1)
      throw std::runtime_error("fused coefficient gather supports one active coefficient per integral");
    auto u = form.coefficients().at(id.coeffs[0]);
    auto dm = u->function_space()->dofmap();
    cf.n_fused = 1;
    cf.fused[0] = {u->x()->array().data(), dm->dev.data(), dm->nd, dm->bs(), 0};
  }
  return cf;
}

template <typename T>
std::unique_ptr<DeviceArray<std::int8_t>>
markers(const FunctionSpace& V, const std::vector<std::reference_wrapper<const DirichletBC<T>>>& bcs)
{
  std::unique_ptr<DeviceArray<std::int8_t>> mk;
  for (auto& bc : bcs)
    if (V.contains(*bc.get().function_space()))
    {
      if (!mk)
      {
        auto im = V.dofmap()->index_map;
        mk = std::make_unique<DeviceArray<std::int8_t>>((std::size_t)V.dofmap()->index_map_bs()
                                                        * (im->size_local() + im->num_ghosts()));
      }
      bc.get().mark_dofs(*mk);
    }
  return mk;
}
} // namespace impl

/// fem::assemble_matrix(A.mat_add_values(), a, bcs) — fem/assembler.h:513-630.  Does not zero A.
template <typename T>
void assemble_matrix(la::MatrixCSR<T>& A, const Form<T>& a,
                     const std::vector<std::reference_wrapper<const DirichletBC<T>>>& bcs = {})
{
  auto mk0 = impl::markers(*a.function_spaces()[0], bcs);
  // one marker array when test and trial space coincide (what the symmetric chunk plan requires)
  const bool one_space = a.function_spaces()[0] == a.function_spaces()[1];
  auto mk1_own = one_space ? nullptr : impl::markers(*a.function_spaces()[1], bcs);
  auto* mk1 = one_space ? mk0.get() : mk1_own.get();
  const std::vector<T> c = pack_constants(a);
  for (auto& [key, id] : a.integrals())
  {
    bfx_asm_t* plan = a.plan(key, A.csr());
    bfx_coeffs_t cf = impl::coeffs_for(a, id);
    // a matrix known to be zero (fresh, or after set(0)) is overwritten by the aggregated kernels instead of added to
    const int mode = A.known_zero() ? BFX_VALUES_OVERWRITE : BFX_VALUES_ADD;
    if (key.first == IntegralType::cell)
      check(bfx_assemble_matrix_cells(plan, id.kernel, a.mesh()->x.data(), mk0 ? mk0->data() : nullptr,
                                      mk1 ? mk1->data() : nullptr, &cf, c.data(), (int)c.size(), A.values().data(),
                                      a.matrix_strategy(key, A.csr()), mode, nullptr));
    else if (key.first == IntegralType::exterior_facet)
    {
      DeviceArray<std::int32_t> ent{std::span<const std::int32_t>(id.entities)};
      check(bfx_assemble_matrix_facets(plan, id.kernel, a.mesh()->x.data(), ent.data(), (std::int64_t)id.entities.size() / 2,
                                       mk0 ? mk0->data() : nullptr, mk1 ? mk1->data() : nullptr, &cf, c.data(),
                                       (int)c.size(), A.values().data(), nullptr));
      check(bfx_stream_sync(nullptr));
    }
    else
      throw std::runtime_error("integral type outside the hot path");
    A.touched();
  }
}

/// fem::assemble_matrix(mat_add, a, bcs) with the matrix's own insertion functor (fem/assembler.h:589-602)
template <typename T>
void assemble_matrix(typename la::MatrixCSR<T>::MatAdd mat_add, const Form<T>& a,
                     const std::vector<std::reference_wrapper<const DirichletBC<T>>>& bcs = {})
{
  assemble_matrix(*mat_add.A, a, bcs);
}

/// fem::assemble_vector(b, L) — fem/assembler.h:230-257
template <typename T>
void assemble_vector(la::Vector<T>& b, const Form<T>& L)
{
  const std::vector<T> c = pack_constants(L);
  for (auto& [key, id] : L.integrals())
  {
    bfx_asm_t* plan = L.plan(key, nullptr);
    bfx_coeffs_t cf = impl::coeffs_for(L, id);
    if (key.first == IntegralType::cell)
      check(bfx_assemble_vector_cells(plan, id.kernel, L.mesh()->x.data(), &cf, c.data(), (int)c.size(),
                                      b.array().data(), BFX_ASM_ATOMIC, nullptr));
    else if (key.first == IntegralType::exterior_facet)
    {
      DeviceArray<std::int32_t> ent{std::span<const std::int32_t>(id.entities)};
      check(bfx_assemble_vector_facets(plan, id.kernel, L.mesh()->x.data(), ent.data(), (std::int64_t)id.entities.size() / 2,
                                       &cf, c.data(), (int)c.size(), b.array().data(), nullptr));
      check(bfx_stream_sync(nullptr));
    }
    else
      throw std::runtime_error("integral type outside the hot path");
  }
}

/// fem::apply_lifting — fem/assembler.h:336-493: b <- b - alpha A_j (g_j - x0_j)
template <typename T>
void apply_lifting(la::Vector<T>& b, const std::vector<std::optional<std::reference_wrapper<const Form<T>>>>& a,
                   const std::vector<std::vector<std::reference_wrapper<const DirichletBC<T>>>>& bcs1,
                   const std::vector<const DeviceArray<T>*>& x0, T alpha)
{
  if (std::ranges::all_of(a, [](auto ai) { return !ai; }))
    return;
  if (!x0.empty() and x0.size() != a.size())
    throw std::runtime_error("Mismatch in size between x0 and bilinear form in assembler.");
  if (a.size() != bcs1.size())
    throw std::runtime_error("Mismatch in size between a and bcs in assembler.");
  for (std::size_t j = 0; j < a.size(); ++j)
  {
    if (!a[j] or bcs1[j].empty())
      continue;
    const Form<T>& aj = a[j]->get();
    auto V1 = aj.function_spaces()[1];
    auto im1 = V1->dofmap()->index_map;
    const std::size_t crange = (std::size_t)V1->dofmap()->index_map_bs() * (im1->size_local() + im1->num_ghosts());
    DeviceArray<std::int8_t> bc_markers1(crange);
    DeviceArray<T> bc_values1(crange);
    for (auto& bc : bcs1[j])
    {
      bc.get().mark_dofs(bc_markers1);
      bc.get().set(bc_values1, nullptr, 1);
    }
    const std::vector<T> c = pack_constants(aj);
    for (auto& [key, id] : aj.integrals())
    {
      if (key.first != IntegralType::cell)
        throw std::runtime_error("lifting of facet integrals is outside the hot path");
      bfx_coeffs_t cf = impl::coeffs_for(aj, id);
      check(bfx_lift_bc_cells(aj.plan(key, nullptr), id.kernel, aj.mesh()->x.data(), &cf, c.data(), (int)c.size(),
                              b.array().data(), bc_values1.data(), bc_markers1.data(),
                              x0.empty() ? nullptr : x0[j]->data(), alpha, nullptr));
    }
    check(bfx_stream_sync(nullptr));
  }
}

/// fem::set_diagonal — fem/assembler.h:644-686 (owned bc rows only, SET)
template <typename T>
void set_diagonal(la::MatrixCSR<T>& A, const FunctionSpace& V,
                  const std::vector<std::reference_wrapper<const DirichletBC<T>>>& bcs, T diagonal = 1.0)
{
  for (auto& bc : bcs)
    if (V.contains(*bc.get().function_space()))
    {
      const auto [dofs, range] = bc.get().dof_indices();
      check(bfx_csr_set_diagonal(A.csr(), A.values().data(), bc.get().dofs_dev(), range, diagonal, nullptr));
    }
}
} // namespace fem
} // namespace dolfinx_b200
