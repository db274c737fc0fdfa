"""The training launcher end to end on the GPU (transoar_b200.train): synthetic cases -> collator -> TrainStep ->
the reference's run formats (runs/<exp>/config.json, model_last.pt), then a resume from that checkpoint."""
import json

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("graph", [False, True])
def test_two_epochs_checkpoint_and_resume(tmp_path, graph):
    from transoar_b200 import train as T
    from transoar_b200.transoarnet import TransoarNet
    prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        base = ["--config", "visceral", "--synthetic", "6", "--volume", "64", "64", "128", "--runs", str(tmp_path), "--deterministic"]
        base += ["--graph"] if graph else []
        assert T.main(base + ["--epochs", "2"]) == 0
        run = tmp_path / "foc_dec_visceral"
        cfg = json.load(open(run / "config.json"))
        assert cfg["neck"]["num_queries"] == 540 and cfg["neck_input_shape"] == [16, 16, 32] and "pytorch_version" in cfg
        assert {"bbox_properties", "loss_coefs", "lr", "lr_backbone", "lr_drop", "epochs", "experiment_name"} <= set(cfg)
        log = [json.loads(l) for l in open(run / "train_log.jsonl")]
        assert [r["epoch"] for r in log] == [1, 2] and all(r["steps_per_rank"] == 3 for r in log)
        assert all(r["train_total_loss"] == r["train_total_loss"] and r["val_total_loss"] == r["val_total_loss"] for r in log)
        assert log[1]["train_total_loss"] < log[0]["train_total_loss"]
        blob = torch.load(run / "model_last.pt", weights_only=False, map_location="cpu")
        assert list(blob) == ["epoch", "metric_max_val", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict"]
        assert blob["epoch"] == 2 and blob["scheduler_state_dict"]["last_epoch"] == 2
        from transoar_b200.engine import visceral_train_config
        ref_cfg = visceral_train_config()
        ref_cfg["neck_input_shape"] = (16, 16, 32)
        fresh = TransoarNet(ref_cfg)
        assert list(blob["model_state_dict"]) == list(fresh.state_dict())                     # the reference's parameter names, in order
        fresh.load_state_dict(blob["model_state_dict"])
        # resume: one more epoch continues the count and the optimiser state
        assert T.main(base + ["--epochs", "3", "--resume", str(run / "model_last.pt")]) == 0
        log = [json.loads(l) for l in open(run / "train_log.jsonl")]
        assert [r["epoch"] for r in log] == [1, 2, 3]
        again = torch.load(run / "model_last.pt", weights_only=False, map_location="cpu")
        assert again["epoch"] == 3
        step = again["optimizer_state_dict"]["state"][0]["step"]
        assert float(step) == 9.0                                                              # 3 epochs x 3 steps, carried over the resume
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = prev
        torch.cuda.empty_cache()
