"""dev tool: one traced engine call (KMCPG_TRACE=1) over the bench's 1 M-read batch in both engine modes"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from kmcp_b200 import api
ctx = api.Context(0)
ctx.build_synth_db(bench.GENOME_SEED, bench.N_GENOMES, bench.GENOME_LEN, k=bench.K, n_chunks=bench.N_CHUNKS, overlap=bench.OVERLAP, num_hashes=bench.H, fpr=bench.FPR,
                   block_size=bench.BLOCK_SIZE)
NR, RL = bench.READS_PER_STEP, bench.READ_LEN
d = ctx.device_alloc(NR * RL)
ctx.synth_reads(bench.READ_SEED, 0, NR, RL, bench.GENOME_SEED, bench.N_GENOMES, bench.GENOME_LEN, d)
pin, ptr = api.pinned_array(NR * RL)
pin[:] = np.frombuffer(ctx.d2h(d, NR * RL), np.uint8)
off = np.arange(NR + 1, dtype=np.uint64) * np.uint64(RL)
eo = ctx.default_engine_opts()
for rep in range(4):
    if rep == 3: sys.stderr.write("==== traced call ====\n")
    if rep < 3:
        fd = os.dup(2); dn = os.open(os.devnull, os.O_WRONLY); os.dup2(dn, 2)
    r = ctx.engine_search_ptr(ptr, off.ctypes.data, NR, eo, copy=False)
    if rep < 3:
        os.dup2(fd, 2); os.close(fd); os.close(dn)
sys.stderr.write("engine ms_total %.2f gpu %.2f post %.2f\n" % (r.ms_total, r.ms_gpu_total, r.ms_post))
