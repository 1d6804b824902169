#!/bin/bash
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run11.log; : > $L
timeout 300 python - >> $L 2>&1 <<'PY'
import torch
from bflc_demo_b200._native import C
from bflc_demo_b200.config import FLConfig
from bflc_demo_b200.data.synthetic import cifar_like
from bflc_demo_b200.engine.generic import GenericFedEngine
from bflc_demo_b200.models.nets import LeNet5
cfg = FLConfig.for_world(1, batch_size=64, samples_per_client=256, learning_rate=0.05, model="lenet5")
eng = GenericFedEngine(cfg, LeNet5(10), cifar_like(1, 256, seed=2, alpha=0.0)[0])
try:
    for _ in range(3): eng.run_round()
except Exception as e:
    print("FAILED", str(e)[:300])
torch.cuda.synchronize()
print("PDL_FALLBACKS", C().pdl_fallbacks(), "launches", C().launch_count())
PY
grep -vE "Warn|warn|^$" $L | tail -c 3000
