"""One (songs, queries, k) case run a few times: the target of `ncu -k regex:scan_kernel --launch-skip N --launch-count 1`."""
import sys
import torch
sys.path.insert(0, ".")
from spotify_recommender_b200 import synth
from spotify_recommender_b200.engine import Engine
n = int(float(sys.argv[1])); nq = int(sys.argv[2]); k = int(sys.argv[3]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
e = Engine(0)
e.load_features(synth.features(n))
for kv in (sys.argv[5] if len(sys.argv) > 5 else "").split(","):
    if "=" in kv:
        e.set_option(kv.split("=")[0], int(kv.split("=")[1]))
q = torch.from_numpy(synth.query_indices(nq, n)).cuda()
oi = torch.empty((nq, k), dtype=torch.int32, device="cuda")
for _ in range(reps):
    e.query_by_index_dev(q, nq, k, oi, None, 0)
torch.cuda.synchronize()
e.synchronize()
print("done", oi[0, :3].tolist())
