"""configs[1] alone: throughput at batch 4096 and batch-1 p50 of the log-barrier kernel (bench.measure_log_variant).
usage: python tools/log_bench.py [K] [repeats]   (SIHMPC_B200_LIB selects an A/B library)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.cuda.set_device(0)
cx = bench.Ctx(0, 1, 0, torch.device("cuda", 0), torch, None)
r = bench.measure_log_variant(cx, K, 3, rep)
print(json.dumps({k: r[k] for k in ("value", "value_min", "value_max", "ms_per_step", "mean_newton_iterations_last_step",
                                    "status_or", "latency_batch1")}))
