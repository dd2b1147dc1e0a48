import sys, time, os, torch, numpy as np
sys.path.insert(0, 'revisit-bpr_b200'); sys.path.insert(0, '.')
import bench
from rbpr import native
from rbpr.engine import Engine
inter = bench.load_interactions('ml-20m', 1.0)
dev = torch.device('cuda:0')
D = int(os.environ.get('D', 128))
ue, ie = bench.init_tables(inter.num_users, inter.num_items, D)
eng = Engine(ue.to(dev), ie.to(dev))
eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
eng.set_reg(bench.REG); eng.set_sgd(0.001); eng.set_sampler(native.SAMPLER_UNIFORM)
perm = torch.randperm(inter.nnz, device=dev)
import itertools
for B, CH in itertools.product((256, 4096, 65536, 262144, 1048576), (0,)):
    os.environ['RBPR_CHUNK'] = str(CH)
    K = min(100, inter.nnz // B)
    t = perm[:K * B].contiguous()
    eng.train_steps(t, B, 1, 0); torch.cuda.synchronize()
    for rep in range(1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        c0 = time.perf_counter(); e0.record()
        eng.train_steps(t, B, 1, 0, want_stats=False)
        e1.record(); c1 = time.perf_counter()
        torch.cuda.synchronize(); c2 = time.perf_counter()
        print(f"chunk={CH} B={B} K={K} cpu_call={1e3*(c1-c0):.2f} ms total={1e3*(c2-c0):.2f} ms gpu={e0.elapsed_time(e1):.2f} ms "
              f"per-step gpu={1e3*e0.elapsed_time(e1)/K:.1f} us  Mtriples/s={K*B/e0.elapsed_time(e1)/1e3:.1f}")
