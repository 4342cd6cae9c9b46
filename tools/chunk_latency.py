"""Fixed cost of the chunk-sharded exchange: tiny batches (local binning ~0) through mtsvgpu_bin_batch_chunked with
MTSV_B200_TRACE=1, one rank per GPU.  torchrun --nproc-per-node N tools/chunk_latency.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mtsv_tools_b200 import MGIndex, Params, synth, chunked  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", lr))
ref = synth.make_reference(8, 50000, seed=100 + rank, n_frac=0.001)
gix = MGIndex.build(ref[0], ref[1], ref[2], ref[3], device=lr)
allref = synth.make_reference(8, 50000, seed=100, n_frac=0.001)
for n in (2000, 200000):
    reads, off = synth.make_reads(allref[0], allref[1], n, 150, seed=7)
    d_r = torch.from_numpy(reads).cuda()
    d_o = torch.from_numpy(off.astype(np.int64)).cuda()
    gix.set_stream(torch.cuda.current_stream().cuda_stream)
    comm = chunked.ChunkComm(lr, max_local_reads=n, max_hits_per_source=4 * n + 1000)
    for it in range(12):
        if it == 6:
            dist.barrier()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        comm.bin_reads(gix, d_r.data_ptr(), d_o.data_ptr(), n, Params())
        dt = (time.perf_counter() - t0) * 1e3
        if it >= 6:
            print("rank %d n=%d call %.3f ms" % (rank, n, dt), file=sys.stderr, flush=True)
    comm.close()
gix.close()
dist.destroy_process_group()
