"""world_size-2 gloo worker for tests/test_sharded.py (run as a subprocess, one per rank)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primitive3d_b200.sharded import exchange_counts_and_tables, exclusive_offsets, gather_counts  # noqa: E402

rank, world, port = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
counts = gather_counts(100 + rank, 1000 + 10 * rank, torch.device("cpu"))
table = torch.arange(24, dtype=torch.int32) + 1000 * rank
counts2, tables = exchange_counts_and_tables(2 ** 33 + rank, 2 ** 40 + 10 * rank, table)
print(json.dumps({"rank": rank, "counts": counts, "offsets": exclusive_offsets(counts, rank), "counts2": counts2,
                  "tables": tables.tolist()}))
dist.barrier()
dist.destroy_process_group()
