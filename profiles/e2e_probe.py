"""profiles/e2e_probe.py - what a host-buffer step costs piece by piece (8192 chains x N=256): the pipelined call with and
without the gather, the sweeps alone, and the individual synchronous entry points.  Run on the GPU box."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/profiles')
smcb = importlib.import_module("montecarlo-surfacer_b200")
N, C, S = 256, 8192, 40
L, Lz, T = 33.0, 240.0, 1.1
nxy = 4; a = L/nxy
cells = np.array([(i, j, k) for i in range(nxy) for j in range(nxy) for k in range(4)], dtype=float)
basis = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
X = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a + a / 4
R0 = X.reshape(-1)
eng = smcb.Engine(C, N, 3)
eng.set_params(smcb.default_params(L=L, Lz=Lz, T=T, A=T), smcb.REFERENCE_WALL_M3)
eng.obs_configure(64, -8.0, 2.0)
eng.broadcast_positions(R0); eng.set_rng(12345, 0, 0)
host_R = torch.empty((C, 3*N), dtype=torch.float64).pin_memory().numpy()
host_R[:] = eng.get_positions()
hE = torch.empty(C, dtype=torch.float64).pin_memory().numpy(); hna = torch.empty(C, dtype=torch.int64).pin_memory().numpy(); hnt = torch.empty(C, dtype=torch.int64).pin_memory().numpy()
for w in range(3): eng.sweep_host(host_R, S, smcb.FAST, gather=True, E=hE, naccept=hna, ntrials=hnt)
for gather in (True, False):
    t0 = time.perf_counter(); ms = []
    for k in range(6):
        eng.sweep_host(host_R, S, smcb.FAST, gather=gather, E=hE, naccept=hna, ntrials=hnt); ms.append(eng.last_kernel_ms()[0])
    wall = (time.perf_counter()-t0)/6*1e3
    print(f"sweep_host gather={gather}: wall {wall:.2f} ms/step, events {np.mean(ms):.2f} ms")
# device-resident reference
ms=[]
for k in range(6):
    eng.sweep(S, smcb.FAST); ms.append(eng.last_kernel_ms()[0])
print(f"sweep only: {np.mean(ms):.2f} ms"); 
t0=time.perf_counter()
for k in range(5): eng.set_positions(host_R)
print(f"set_positions {(time.perf_counter()-t0)/5*1e3:.2f} ms")
t0=time.perf_counter()
for k in range(5): eng.get_positions()
print(f"get_positions(pageable) {(time.perf_counter()-t0)/5*1e3:.2f} ms")
t0=time.perf_counter()
for k in range(5): eng.refresh_energy(smcb.FAST)
print(f"refresh_energy {(time.perf_counter()-t0)/5*1e3:.2f} ms")
t0=time.perf_counter()
for k in range(5): eng.gather()
print(f"gather {(time.perf_counter()-t0)/5*1e3:.2f} ms")
