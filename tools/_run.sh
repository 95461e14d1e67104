mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cahn_gpu.py -q -m gpu -x -k "last_pass or steps_compose or graph_replay or switching or two_solvers" 2>&1 | tail -12
for f in 1 0; do python - <<P
import numpy as np, custen_b200 as cs
from custen_b200.cahn import CahnHilliard
cs.load().custen_cahn_set_fuse_new($f)
for n in (4096, 2048, 512):
    s = CahnHilliard(n, solver=2); s.set_field(np.random.default_rng(0).uniform(-0.1, 0.1, (n, n))); s.step(8)
    print("fuse_new", $f, "n", n, "ms/step", [round(s.time_steps(40), 4) for _ in range(3)]); s.destroy()
P
done
