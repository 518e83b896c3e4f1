"""world_size-2 gloo worker for tests/test_verify.py: each rank holds one virtual shard of an oracle mesh (npz written
by the test) and enters primitive3d_b200.verify.mesh_checksums."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primitive3d_b200.verify import mesh_checksums  # noqa: E402

rank, world, port, path = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
z = np.load(path)
v, f = torch.from_numpy(z[f"v{rank}"]), torch.from_numpy(z[f"f{rank}"])
sums = mesh_checksums(v, f, int(z["voff"][rank]), int(z["foff"][rank]), float(z["plane"][rank]), chunk=1000)
print(json.dumps({"rank": rank, "sums": list(sums)}))
dist.barrier()
dist.destroy_process_group()
