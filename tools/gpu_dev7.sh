#!/bin/bash
mkdir -p gpurun_out/dev7
python - > gpurun_out/dev7/profile.txt 2>&1 <<'PY'
import cProfile, pstats, io, os, sys, time, torch
sys.path.insert(0, os.getcwd())
import equilib_b200
dev = torch.device("cuda", 0)
equi = torch.rand((3, 2048, 4096), device=dev)
rot = {"roll": 0.1, "pitch": 0.3, "yaw": -0.5}
op = equilib_b200.Equi2Pers(480, 640, 90.0)
for _ in range(50): op(equi, rot)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000): op(equi, rot)
torch.cuda.synchronize()
print("wall per call us:", (time.perf_counter() - t0) / 2000 * 1e6)
pr = cProfile.Profile()
pr.enable()
for _ in range(2000): op(equi, rot)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35)
print(s.getvalue())
PY
cat gpurun_out/dev7/profile.txt | head -80
