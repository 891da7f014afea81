// Row-partitioned levels: halo exchange (peer memory / NCCL), the partitioned cycle and its graph.  Part of engine.cu.
#pragma once

// ------------------------------------------------------------------------------------------
// Row-partitioned fine level (config C4): halo exchange over NCCL, coarse levels on rank 0
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) halo_pack_kernel(int n, const int* __restrict__ idx, const double* __restrict__ v,
                                                             double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v[idx[i]];
}

// v is laid out [owned | halo]: gather what the neighbours need, exchange, receive straight into the halo
static int peer_channel(const Part& P, const double* v) { return v == P.x ? 0 : v == P.temp ? 1 : v == P.res ? 2 : -1; }
// the kernel(s) that read the halo of the last exchange have been enqueued on the compute stream: tell the senders
static void halo_consumed(H* h) {
  if (!h->peer_pending) return;
  const int slot = h->peer_pending->level * kPeerChannels + h->peer_pending_ch;
  halo_ack_kernel<<<1, 32, 0, h->stream>>>(h->peer.d_tab + slot, h->peer.sync + (size_t)slot * kPeerWords);
  count_launch(h);
  h->peer_pending = nullptr;
}
static void halo_exchange(H* h, Part& P, double* v) {
  const PartPlan& pl = P.plan;
  if (h->peer.on) {
    const int ch = peer_channel(P, v);
    REQUIRE(ch >= 0, B200AMG_ERR_STATE, "peer halo exchange of a vector that was not exported");
    REQUIRE(!h->peer_pending, B200AMG_ERR_STATE, "internal: a halo exchange was started before the previous one was acknowledged");
    const int slot = P.level * kPeerChannels + ch;
    unsigned long long* sync = h->peer.sync + (size_t)slot * kPeerWords;
    const int nsend = pl.send_off[pl.world];
    const int blocks = std::max(1, std::min(64, (nsend + 255) / 256));
    halo_push_kernel<<<blocks, 256, 0, h->stream>>>(h->peer.d_tab + slot, P.send_idx, v, sync, h->peer.tickets + slot, h->gs_fault);
    count_launch(h);
    halo_wait_kernel<<<1, 32, 0, h->stream>>>(h->peer.d_tab + slot, sync, h->gs_fault);
    count_launch(h);
    h->peer_pending = &P;
    h->peer_pending_ch = ch;
    h->peer_exchanges++;
    return;
  }
  NcclApi& nc = nccl_api();
  const int nsend = pl.send_off[pl.world];
  if (nsend > 0) {
    halo_pack_kernel<<<grid_for(nsend), kThreads, 0, h->stream>>>(nsend, P.send_idx, v, P.sendbuf);
    count_launch(h);
  }
  NCCL_OK(nc.GroupStart());
  for (int q = 0; q < pl.world; ++q) {
    if (q == pl.rank) continue;
    const int ns = pl.send_off[q + 1] - pl.send_off[q], nr = pl.recv_off[q + 1] - pl.recv_off[q];
    if (ns > 0) NCCL_OK(nc.Send(P.sendbuf + pl.send_off[q], (size_t)ns, ncclDouble, q, h->comm, h->stream));
    if (nr > 0) NCCL_OK(nc.Recv(v + pl.nloc + pl.recv_off[q], (size_t)nr, ncclDouble, q, h->comm, h->stream));
  }
  NCCL_OK(nc.GroupEnd());
  h->collectives++;
}

// The same exchange, split in two so that work which does not touch the halo can run in between: _begin forks onto the
// communication stream (after everything enqueued so far on the compute stream: producers of v's owned part, earlier readers
// of its halo part), _end joins it back.  Both are captured into the whole-cycle graph as a fork / join.
static void halo_exchange_begin(H* h, Part& P, double* v) {
  if (!h->part_overlap) { halo_exchange(h, P, v); return; }
  cudaStream_t compute = h->stream;
  CUDA_OK(cudaEventRecord(h->ev_ready, compute));
  CUDA_OK(cudaStreamWaitEvent(h->comm_stream, h->ev_ready, 0));
  h->stream = h->comm_stream;
  try {
    halo_exchange(h, P, v);
  } catch (...) {
    h->stream = compute;
    throw;
  }
  h->stream = compute;
  CUDA_OK(cudaEventRecord(h->ev_done, h->comm_stream));
}
static void halo_exchange_end(H* h) {
  if (!h->part_overlap) return;
  CUDA_OK(cudaStreamWaitEvent(h->stream, h->ev_done, 0));
}

static void smooth_part(H* h, Part& P, const SmootherCfg& c) {
  if (c.kind == B200AMG_SMOOTHER_NONE || P.plan.nloc == 0) {
    if (c.kind != B200AMG_SMOOTHER_NONE)
      for (int it = 0; it < c.iter; ++it) { halo_exchange(h, P, P.x); halo_consumed(h); }   // keep the exchanges matched
    return;
  }
  double* cur = P.x;
  double* other = P.temp;
  for (int it = 0; it < c.iter; ++it) {
    halo_exchange_begin(h, P, cur);
    for (int part = 1; part <= 2; ++part) {   // interior rows while the halo is in flight, boundary rows after it has landed
      if (part == 2) halo_exchange_end(h);
      const int sel = h->part_overlap ? part : (part == 2 ? 0 : -1);
      if (sel < 0) continue;
      if (P.symmetry == B200AMG_SYMMETRY_NONE) launch_jacobi_general(h, P.A, P.diag, cur, P.b, other, c.omega, sel);
      else launch_jacobi_fast(h, P.walked(), cur, P.b, other, c.omega, sel);
    }
    halo_consumed(h);
    std::swap(cur, other);
  }
  if (cur != P.x) CUDA_OK(cudaMemcpyAsync(P.x, cur, sizeof(double) * (size_t)P.plan.nloc, cudaMemcpyDeviceToDevice, h->stream));
}

static void solve_level(H* h, double* x, int cycle, const double* b, int lvl, bool x_is_zero);
static void coarse_solve(H* h, double* x, const double* b);

// __solve!(x, ml, cycle, b, lvl) for a level split by rows (multilevel.jl:214-239)
static void cycle_part_level(H* h, int lvl, int cycle) {
  Part& P = *h->parts[lvl];
  const PartPlan& pl = P.plan;
  NcclApi& nc = nccl_api();
  Level& L0 = *h->levels[lvl];
  smooth_part(h, P, P.pre);                                                      // :216
  auto split = [&](auto&& launch) {   // interior part, join the exchange, boundary part (or everything after a blocking exchange)
    if (h->part_overlap) { launch(1); halo_exchange_end(h); launch(2); }
    else launch(0);
    halo_consumed(h);
  };
  halo_exchange_begin(h, P, P.x);
  split([&](int part) { residual(h, P.A, P.x, P.b, P.res, part); });               // :219-220
  halo_exchange_begin(h, P, P.res);
  split([&](int part) { spmv(h, P.R, P.res, P.cb, part); });                       // :223 (my coarse rows)
  if (P.child) {
    // the level below is partitioned too: the restriction wrote straight into its b (the rows I own there)
    Part& C = *P.child;
    CUDA_OK(cudaMemsetAsync(C.x, 0, sizeof(double) * (size_t)std::max<int64_t>(C.plan.nloc, 1), h->stream));   // :226
    if (cycle == B200AMG_CYCLE_V) {
      cycle_part_level(h, lvl + 1, B200AMG_CYCLE_V);
    } else if (cycle == B200AMG_CYCLE_W) {
      cycle_part_level(h, lvl + 1, B200AMG_CYCLE_W);
      cycle_part_level(h, lvl + 1, B200AMG_CYCLE_W);
    } else {
      cycle_part_level(h, lvl + 1, B200AMG_CYCLE_F);
      cycle_part_level(h, lvl + 1, B200AMG_CYCLE_V);
    }
    halo_exchange_begin(h, C, C.x);                                              // my rows of P reach into the neighbours' coarse entries
    split([&](int part) { spmv_add(h, P.P, C.x, P.x, part); });                    // :233-234
    smooth_part(h, P, P.post);                                                   // :236
    return;
  }
  NCCL_OK(nc.GroupStart());                                                      // coarse_b -> rank 0
  if (pl.rank == 0) {
    for (int q = 1; q < pl.world; ++q) {
      const int64_t cnt = pl.coarse_split[q + 1] - pl.coarse_split[q];
      if (cnt > 0) NCCL_OK(nc.Recv(L0.coarse_b + pl.coarse_split[q], (size_t)cnt, ncclDouble, q, h->comm, h->stream));
    }
  } else if (pl.ncloc > 0) {
    NCCL_OK(nc.Send(P.cb, (size_t)pl.ncloc, ncclDouble, 0, h->comm, h->stream));
  }
  NCCL_OK(nc.GroupEnd());
  h->collectives++;
  if (pl.rank == 0) {
    // everything below the partitioned level is a static kernel sequence on this rank: one graph per cycle type
    auto coarse_part = [&]() {
      CUDA_OK(cudaMemsetAsync(L0.coarse_x, 0, sizeof(double) * (size_t)std::max<int64_t>(L0.nc, 1), h->stream));   // :226
      if ((int)h->levels.size() == lvl + 1) {
        coarse_solve(h, L0.coarse_x, L0.coarse_b);                                  // :228
      } else if (cycle == B200AMG_CYCLE_V) {
        solve_level(h, L0.coarse_x, B200AMG_CYCLE_V, L0.coarse_b, lvl + 1, true);
      } else if (cycle == B200AMG_CYCLE_W) {
        solve_level(h, L0.coarse_x, B200AMG_CYCLE_W, L0.coarse_b, lvl + 1, true);
        solve_level(h, L0.coarse_x, B200AMG_CYCLE_W, L0.coarse_b, lvl + 1, false);
      } else {
        solve_level(h, L0.coarse_x, B200AMG_CYCLE_F, L0.coarse_b, lvl + 1, true);
        solve_level(h, L0.coarse_x, B200AMG_CYCLE_V, L0.coarse_b, lvl + 1, false);
      }
    };
    if (h->part_graphs && !h->part_whole_graph && !h->capturing && !h->cycle_graph[cycle] && h->cycle_graph_launches[cycle] >= 0 && !h->profiling) {
      cudaGraph_t gr = nullptr;
      h->capturing = true;
      h->capture_count = 0;
      CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
      try {
        coarse_part();
      } catch (...) {
        cudaStreamEndCapture(h->stream, &gr);
        if (gr) cudaGraphDestroy(gr);
        h->capturing = false;
        throw;
      }
      CUDA_OK(cudaStreamEndCapture(h->stream, &gr));
      h->capturing = false;
      CUDA_OK(cudaGraphInstantiate(&h->cycle_graph[cycle], gr, 0));
      CUDA_OK(cudaGraphDestroy(gr));
      h->cycle_graph_launches[cycle] = h->capture_count;
    }
    if (h->part_graphs && !h->part_whole_graph && !h->capturing && h->cycle_graph[cycle]) {
      CUDA_OK(cudaGraphLaunch(h->cycle_graph[cycle], h->stream));
      h->launches += h->cycle_graph_launches[cycle];
    } else {
      coarse_part();
    }
  }
  NCCL_OK(nc.GroupStart());                                                      // coarse_x windows <- rank 0
  if (pl.rank == 0) {
    for (int q = 1; q < pl.world; ++q) {
      const int64_t cnt = pl.cx_hi_all[q] - pl.cx_lo_all[q];
      if (cnt > 0) NCCL_OK(nc.Send(L0.coarse_x + pl.cx_lo_all[q], (size_t)cnt, ncclDouble, q, h->comm, h->stream));
    }
  } else if (pl.cx_hi > pl.cx_lo) {
    NCCL_OK(nc.Recv(P.cx, (size_t)(pl.cx_hi - pl.cx_lo), ncclDouble, 0, h->comm, h->stream));
  }
  NCCL_OK(nc.GroupEnd());
  h->collectives++;
  spmv_add(h, P.P, P.cx, P.x);                                                   // :233-234
  smooth_part(h, P, P.post);                                                     // :236
}
// The whole partitioned cycle — kernels, memsets and the NCCL point-to-point groups of every level — is a static sequence on
// every rank, so it is captured ONCE per cycle type into a CUDA graph and replayed (NCCL >= 2.9 records its kernels into a
// capturing stream): ~60 launches + ~18 communication groups per V-cycle become one graph launch per rank.
// B200AMG_PART_WHOLE_GRAPH=0 (or USE_GRAPHS=0 after finalize) goes back to eager launches.
static void cycle_body_part(H* h, int cycle) {
  if (!h->part_whole_graph || h->profiling) { cycle_part_level(h, 0, cycle); return; }
  if (!h->part_cycle_graph[cycle]) {
    cudaGraph_t g = nullptr;
    const int64_t coll0 = h->collectives, peer0 = h->peer_exchanges;
    h->capturing = true;
    h->capture_count = 0;
    CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    try {
      cycle_part_level(h, 0, cycle);
    } catch (...) {
      cudaStreamEndCapture(h->stream, &g);
      if (g) cudaGraphDestroy(g);
      h->capturing = false;
      throw;
    }
    CUDA_OK(cudaStreamEndCapture(h->stream, &g));
    h->capturing = false;
    CUDA_OK(cudaGraphInstantiate(&h->part_cycle_graph[cycle], g, 0));
    CUDA_OK(cudaGraphDestroy(g));
    h->part_cycle_launches[cycle] = h->capture_count;
    h->part_cycle_collectives[cycle] = h->collectives - coll0;
    h->collectives = coll0;
    h->part_cycle_peer[cycle] = h->peer_exchanges - peer0;
    h->peer_exchanges = peer0;
  }
  CUDA_OK(cudaGraphLaunch(h->part_cycle_graph[cycle], h->stream));
  h->launches += h->part_cycle_launches[cycle];
  h->collectives += h->part_cycle_collectives[cycle];
  h->peer_exchanges += h->part_cycle_peer[cycle];
}

// sum over ranks of a device scalar, in place; every rank gets the same bits
static void allreduce_scalar(H* h, double* dev) {
  NCCL_OK(nccl_api().AllReduce(dev, dev, 1, ncclDouble, ncclSum, h->comm, h->stream));
  h->collectives++;
}
// scalars[slot] = sum over all ranks of v.v over the owned entries (NOT square-rooted)
static void sumsq_part(H* h, const double* v, double* out_dev) {
  dot_partial_kernel<<<kRedBlocks, kThreads, 0, h->stream>>>(h->part->plan.nloc, v, v, h->partial);
  count_launch(h);
  reduce_final_kernel<<<1, kThreads, 0, h->stream>>>(kRedBlocks, h->partial, out_dev, 0);
  count_launch(h);
  allreduce_scalar(h, out_dev);
}
static void residual_norm_part(H* h) {   // scalars[0] = ||b - A x||^2 over all ranks
  Part& P = *h->part;
  halo_exchange(h, P, P.x);   // (blocking here: the kernel below is the one bench.py times on its own)
  const bool timed = h->time_residual && h->res_events_used + 2 <= (int)h->res_events.size();
  if (timed) CUDA_OK(cudaEventRecord(h->res_events[h->res_events_used], h->stream));
  residual(h, P.A, P.x, P.b, P.res);
  if (timed) {
    CUDA_OK(cudaEventRecord(h->res_events[h->res_events_used + 1], h->stream));
    h->res_events_used += 2;
  }
  halo_consumed(h);
  sumsq_part(h, P.res, h->scalars);
}
// owned slice in, assembled vector out
static void part_load(H* h, double* dst, const double* src_full, int memkind) {
  const PartPlan& pl = h->part->plan;
  if (pl.nloc == 0) return;
  CUDA_OK(cudaMemcpyAsync(dst, src_full + pl.row_split[pl.rank], sizeof(double) * (size_t)pl.nloc,
                          memkind == B200AMG_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, h->stream));
}
static void part_store(H* h, double* dst_full, const double* src_local, int memkind) {
  Part& P = *h->part;
  const PartPlan& pl = P.plan;
  NcclApi& nc = nccl_api();
  double* full = dst_full;
  if (memkind == B200AMG_MEM_HOST) {
    if (!P.xfull) P.xfull = dev_alloc<double>(P.n);
    full = P.xfull;
  }
  NCCL_OK(nc.GroupStart());
  for (int q = 0; q < pl.world; ++q) {
    const int64_t cnt = pl.row_split[q + 1] - pl.row_split[q];
    if (cnt > 0) NCCL_OK(nc.Broadcast(q == pl.rank ? src_local : full + pl.row_split[q], full + pl.row_split[q], (size_t)cnt, ncclDouble, q, h->comm, h->stream));
  }
  NCCL_OK(nc.GroupEnd());
  h->collectives++;
  if (memkind == B200AMG_MEM_HOST) CUDA_OK(cudaMemcpyAsync(dst_full, full, sizeof(double) * (size_t)P.n, cudaMemcpyDeviceToHost, h->stream));
}

