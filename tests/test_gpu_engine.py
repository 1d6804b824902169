"""Single-GPU end-to-end: fused engine rounds, host-ledger re-execution, learning progress,
parity with the NCCL/cuBLAS baseline engine (world = 1, solo mode)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_engine_solo_rounds_learn_and_chain():
    from bflc_demo_b200.config import FLConfig
    from bflc_demo_b200.data.synthetic import femnist_like
    from bflc_demo_b200.engine.fused import FusedEngine
    cfg = FLConfig.for_world(1, hidden=256, batch_size=128, samples_per_client=1024,
                             learning_rate=0.05)
    shard = femnist_like(1, 1024, seed=3)[0]
    eng = FusedEngine(cfg, shard, rank=0, world=1, device=0)
    acc0 = eng.evaluate(shard)
    eng.capture()
    losses = []
    for _ in range(8):
        losses.append(eng.run_round_e2e()["global_loss"])
    assert eng.drain_blocks() == []
    assert eng.host_ledger.n_blocks() == 9 and eng.host_ledger.verify_chain()
    assert losses[-1] < losses[0]
    assert eng.evaluate(shard) > acc0 + 0.2
    blk = eng.host_ledger.blocks()[-1]
    assert blk["from_device"] and blk["selected"] == [0] and blk["device_digest"] != 0


def test_fused_matches_nccl_baseline_one_round():
    from bflc_demo_b200.config import FLConfig
    from bflc_demo_b200.data.synthetic import femnist_like
    from bflc_demo_b200.engine.fused import FusedEngine
    from bflc_demo_b200.engine.nccl_baseline import NcclBaselineEngine
    cfg = FLConfig.for_world(1, hidden=256, batch_size=128, samples_per_client=512,
                             learning_rate=0.05, cuda_graph=False)
    shard = femnist_like(1, 512, seed=5)[0]
    a = FusedEngine(cfg, shard, rank=0, world=1, device=0)
    b = NcclBaselineEngine(cfg, shard, rank=0, world=1, device=0)
    a.run_round()
    torch.cuda.synchronize()
    b.x_bf.copy_(b.x_u8.to(torch.bfloat16) * (1.0 / 255.0))
    rb = b.run_round()
    sa = a.read_state()
    assert sa["epoch"] == rb["epoch"] == 1
    assert abs(sa["global_loss"] - rb["global_loss"]) < 2e-2 * max(1.0, rb["global_loss"])
    wa, wb = a.global_master, b.global_w
    assert ((wa - wb).norm() / wb.norm()).item() < 2e-2
