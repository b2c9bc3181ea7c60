// Host side of several ranks in the C++ mirror (dolfinx_b200.h): IndexMap (src / dest / offsets), the Scatterer plan
// (common/Scatterer.h:65-198), SparsityPattern::finalize with its ghost-row exchange (la/SparsityPattern.cpp:264-491)
// and the MatrixCSR ghost-row plan (la/MatrixCSR.h:705-849), on R simulated ranks = R threads of this process whose
// Comm callbacks exchange through in-memory mailboxes (what MPI_Neighbor_alltoallv / MPI_Allgatherv do in the
// reference).  Input (argv[1]) and output (argv[2]) are whitespace-separated integer files; tests/test_cpp_api.py
// writes the input from a brick partition and compares the output with the oracle bit for bit.  Without a third
// argument no device is touched.  With "device" (needs one GPU per rank): every thread binds its GPU, joins the NCCL
// communicator and runs the data path on those plans - la::Vector scatter_fwd / scatter_rev and
// MatrixCSR::scatter_rev + squared_norm - against values known in closed form.
#include "../../dolfinx_b200/cpp/dolfinx_b200.h"
#include <barrier>
#include <cstdio>
#include <fstream>
#include <mutex>
#include <sstream>
#include <thread>

using namespace dolfinx_b200;
using Buffers = common::Comm::Buffers;

struct World
{
  explicit World(int n) : size(n), sync(n), box(n, Buffers(n)), gathered(n) {}
  int size;
  char nccl_id[128] = {};
  std::barrier<> sync;
  std::vector<Buffers> box; // box[from][to]
  Buffers gathered;

  common::Comm comm(int rank)
  {
    common::Comm c;
    c.rank = rank, c.size = size;
    c.neighbor_alltoallv = [this, rank](std::span<const int> dest, std::span<const int> src, const Buffers& send)
    {
      for (std::size_t i = 0; i < dest.size(); ++i)
        box[rank][dest[i]] = send[i];
      sync.arrive_and_wait();
      Buffers recv;
      for (int s : src)
        recv.push_back(box[s][rank]);
      sync.arrive_and_wait();
      for (std::size_t i = 0; i < dest.size(); ++i)
        box[rank][dest[i]].clear();
      sync.arrive_and_wait();
      return recv;
    };
    c.allgatherv = [this, rank](std::span<const std::int64_t> mine)
    {
      gathered[rank].assign(mine.begin(), mine.end());
      sync.arrive_and_wait();
      Buffers all = gathered;
      sync.arrive_and_wait();
      return all;
    };
    return c;
  }
};

struct RankInput
{
  std::int32_t n_owned = 0;
  std::vector<std::int64_t> ghosts;
  std::vector<int> owners;
  std::vector<std::int32_t> rows, cols;
};

template <class V>
static void put(std::ostream& os, const char* name, const V& v)
{
  os << name << ' ' << v.size();
  for (auto e : v)
    os << ' ' << (long long)e;
  os << '\n';
}

int main(int argc, char** argv)
{
  if (argc < 3)
    return 2;
  std::ifstream in(argv[1]);
  int R = 0, bs = 1;
  in >> R >> bs;
  std::vector<RankInput> inputs(R);
  for (auto& q : inputs)
  {
    std::size_t ng = 0, nnz = 0;
    in >> q.n_owned >> ng;
    q.ghosts.resize(ng), q.owners.resize(ng);
    for (auto& g : q.ghosts)
      in >> g;
    for (auto& o : q.owners)
      in >> o;
    in >> nnz;
    q.rows.resize(nnz), q.cols.resize(nnz);
    for (auto& r : q.rows)
      in >> r;
    for (auto& c : q.cols)
      in >> c;
  }
  if (!in)
    return 3;
  World world(R);
  const bool device = argc > 3 and std::string(argv[3]) == "device";
  std::vector<std::string> out(R);
  std::vector<std::string> errors(R);
  auto work = [&](int rank)
  {
    try
    {
      const RankInput& q = inputs[rank];
      common::Comm comm = world.comm(rank);
      if (device)
      {
        check(bfx_set_device(rank));
        if (rank == 0)
          check(bfx_comm_unique_id(world.nccl_id));
        world.sync.arrive_and_wait();
        check(bfx_comm_create(&comm.nccl, world.nccl_id, rank, R));
      }
      auto map = std::make_shared<const common::IndexMap>(comm, q.n_owned, q.ghosts, q.owners);
      common::Scatterer sc(*map, bs);
      la::SparsityPattern sp(comm, {map, map}, {bs, bs});
      // cell-wise insertion order is what the ghost-row exchange preserves: keep the input order
      for (std::size_t k = 0; k < q.rows.size(); ++k)
        sp.insert(q.rows[k], q.cols[k]);
      sp.finalize();
      auto [edges, offsets] = sp.graph();
      auto cmap = sp.index_map(1);
      const la::impl::GhostRowPlan g = la::impl::matrix_ghost_plan(*map, *cmap, {bs, bs}, offsets, edges);
      std::ostringstream os;
      os << "rank " << rank << '\n';
      const std::array<std::int64_t, 2> lr = map->local_range();
      os << "range " << lr[0] << ' ' << lr[1] << ' ' << map->size_global() << '\n';
      put(os, "src", map->src());
      put(os, "dest", map->dest());
      put(os, "local_inds", sc.local_indices());
      put(os, "remote_inds", sc.remote_indices());
      put(os, "sizes_local", sc.sizes_local());
      put(os, "displs_local", sc.displs_local());
      put(os, "sizes_remote", sc.sizes_remote());
      put(os, "displs_remote", sc.displs_remote());
      put(os, "edges", edges);
      put(os, "offsets", offsets);
      put(os, "off_diag", sp.off_diagonal_offsets());
      put(os, "col_ghosts", cmap->ghosts());
      put(os, "col_owners", cmap->owners());
      put(os, "col_src", cmap->src());
      put(os, "col_dest", cmap->dest());
      put(os, "ghost_row_to_rank", g.ghost_row_to_rank);
      put(os, "val_send_disp", g.val_send_disp);
      put(os, "val_recv_disp", g.val_recv_disp);
      put(os, "unpack_pos", g.unpack_pos);
      out[rank] = os.str();
      if (device)
      {
        // ---- la::Vector: owners -> ghosts, then ghosts -> owners (add)
        auto scp = std::make_shared<const common::Scatterer>(*map, bs);
        la::Vector<double> v(map, bs, scp);
        const std::int32_t nl = map->size_local(), ng = map->num_ghosts();
        std::vector<double> h((std::size_t)bs * (nl + ng), -1.0);
        for (std::int32_t i = 0; i < nl; ++i)
          for (int c = 0; c < bs; ++c)
            h[(std::size_t)bs * i + c] = (double)((lr[0] + i) * bs + c);
        v.array().copy_from(h);
        v.scatter_fwd();
        h = v.array().to_host();
        for (std::int32_t i = 0; i < ng; ++i)
          for (int c = 0; c < bs; ++c)
            if (h[(std::size_t)bs * (nl + i) + c] != (double)(map->ghosts()[i] * bs + c))
              throw std::runtime_error("scatter_fwd: wrong ghost value");
        v.set(1.0);
        v.scatter_rev(std::plus<double>{});
        h = v.array().to_host();
        double owned_sum = 0;
        for (std::int32_t i = 0; i < bs * nl; ++i)
          owned_sum += h[i];
        // every ghost copy adds 1 to its owner: sum over all owned entries = bs (global size + all ghosts)
        const Buffers ghosts_all = comm.allgatherv(std::vector<std::int64_t>{(std::int64_t)ng, (std::int64_t)owned_sum});
        std::int64_t tot_ghosts = 0, tot_sum = 0;
        for (auto& b : ghosts_all)
          tot_ghosts += b[0], tot_sum += b[1];
        if (tot_sum != (std::int64_t)bs * (map->size_global() + tot_ghosts))
          throw std::runtime_error("scatter_rev(add): wrong sum over the owned entries");
        // ---- MatrixCSR: every stored value 1, ghost rows to their owners; |A|_F^2 over the ranks
        la::MatrixCSR<double> A(sp);
        A.set(1.0);
        A.scatter_rev();
        std::vector<std::int64_t> hits(A.cols().size(), 0);
        for (std::int64_t pos : g.unpack_pos)
          hits[pos] += 1;
        std::int64_t expect = 0;
        for (std::int64_t j = 0; j < A.row_ptr()[nl]; ++j)
          expect += (1 + hits[j]) * (1 + hits[j]) * bs * bs;
        const Buffers ex_all = comm.allgatherv(std::vector<std::int64_t>{expect});
        std::int64_t expect_all = 0;
        for (auto& b : ex_all)
          expect_all += b[0];
        const double n2 = A.squared_norm();
        if (n2 != (double)expect_all)
          throw std::runtime_error("MatrixCSR::scatter_rev / squared_norm: " + std::to_string(n2) + " != " + std::to_string(expect_all));
        auto vals = A.values().to_host();
        for (std::size_t k = (std::size_t)A.row_ptr()[nl] * bs * bs; k < vals.size(); ++k)
          if (vals[k] != 0.0)
            throw std::runtime_error("ghost rows must be zero after scatter_rev (la/MatrixCSR.h:462-466)");
        if (rank == 0)
          std::printf("device path on %d ranks: |A|_F^2 = %.0f\n", R, n2);
      }
    }
    catch (const std::exception& e)
    {
      errors[rank] = e.what();
      std::fprintf(stderr, "rank %d: %s\n", rank, e.what());
      std::_Exit(4); // (the other threads wait in a barrier)
    }
  };
  std::vector<std::thread> threads;
  for (int r = 0; r < R; ++r)
    threads.emplace_back(work, r);
  for (auto& t : threads)
    t.join();
  std::ofstream o(argv[2]);
  for (auto& sres : out)
    o << sres;
  std::printf("CPP_MULTIRANK_OK\n");
  return 0;
}
