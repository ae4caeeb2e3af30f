"""Data-parallel training step over NCCL (rows shard, gradients all-reduce): needs >= 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_ddp.py -m gpu`); skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["ddp", "overlap", "overlap32"])
def test_ddp_gradients_match_whole_batch(mode):
    """mode "ddp": the reference's setup, DDP's bucketed all-reduce after the native backward; mode "overlap":
    comm.enable_overlapped_grad_sync — NCCL all-reduce per gradient group from inside the (graph-captured) backward,
    exchanging bf16 copies ("overlap") or the fp32 gradients themselves ("overlap32"); also no_sync() accumulation."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", {"ddp": "29517", "overlap": "29518", "overlap32": "29519"}[mode],
           os.path.join(ROOT, "tests", "helpers", "ddp_train_worker.py"), mode]
    env = dict(os.environ, NCCL_DEBUG="WARN")
    try:
        out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    except subprocess.TimeoutExpired as e:
        def txt(x):
            return x.decode(errors="replace") if isinstance(x, bytes) else (x or "")
        pytest.fail("worker hung; stdout so far:\n" + txt(e.stdout)[-3000:] + "\nstderr:\n" + txt(e.stderr)[-2000:])
    assert out.returncode == 0 and "DDP_TRAIN_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_replicas_get_their_own_handles():
    """torch.nn.DataParallel (gqa_cpt.py:358-359, the non-distributed multi-GPU branch): module replicas share the
    Python object graph but run on other devices from other threads — each device must get its own native handle."""
    from cpt_b200 import config as C
    from cpt_b200.modeling_bert import BertImgForPreTraining
    from cpt_b200.modeling_rec import REC_MLM_CPT
    from cpt_b200.synthetic import synth_batch, synth_state_dict
    cfg = C.oscar_tiny(num_hidden_layers=2)
    sd = synth_state_dict(cfg, seed=8)
    pre = BertImgForPreTraining(cfg)
    pre.load_state_dict(sd, strict=False)
    pre.tie_weights()
    rec = REC_MLM_CPT(cfg)
    rec.copy_from_pretraining_model(pre.cuda(0))
    rec.eval()
    b = synth_batch(cfg, 6, 30, 10, seed=8)
    d = {k: v.cuda(0) for k, v in b.items()}
    with torch.no_grad():
        single = rec(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"])[0]
        dp = torch.nn.DataParallel(rec, device_ids=[0, 1])
        for _ in range(2):
            multi = dp(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"])[0]
    assert multi.shape == single.shape
    assert (multi - single).abs().max().item() <= 1e-5 * single.abs().max().item()
    assert len(rec.bert._slot._per_device) == 1


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_memory_logits_exchange():
    """comm.LogitsExchange: the decoder kernel stores its logits into every rank's gather buffer over NVLink (CUDA IPC
    peer memory) and a flag kernel completes the all-gather — bit-identical to NCCL's all_gather of the same logits,
    through eager, captured and replayed steps."""
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % n, "--master-addr",
           "127.0.0.1", "--master-port", "29521", os.path.join(ROOT, "tests", "helpers", "exchange_worker.py")]
    try:
        out = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, NCCL_DEBUG="WARN"), capture_output=True, text=True,
                             timeout=240)
    except subprocess.TimeoutExpired as e:
        def txt(x):
            return x.decode(errors="replace") if isinstance(x, bytes) else (x or "")
        pytest.fail("worker hung; stdout so far:\n" + txt(e.stdout)[-3000:] + "\nstderr:\n" + txt(e.stderr)[-2000:])
    assert out.returncode == 0 and "EXCHANGE_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
