"""The training entry point keeps the reference's command line, CSV columns and checkpoint convention
(/root/reference/train.py:20-56,167-171,199-213)."""
import csv
import os

import pytest
import torch

import train


def test_cli_flags_and_defaults_match_the_reference():
    a = train.get_args([])
    assert (a.log_dir, a.ckpt_dir, a.dataset, a.pos_encoding) == ("logs", "checkpoints", "mnist", "absolute")
    assert (a.rope_theta, a.poly_degree, a.poly_shared_heads) == (100.0, 3, True)
    assert (a.batch_size, a.epochs, a.lr, a.weight_decay) == (128, 25, 0.001, 0.01)
    assert (a.img_size, a.patch_size, a.embed_dim, a.depth, a.num_heads) == (32, 4, 192, 6, 6)
    assert train.get_args(["--no-poly_shared_heads"]).poly_shared_heads is False
    with pytest.raises(SystemExit):
        train.get_args(["--pos_encoding", "sinusoidal"])
    with pytest.raises(SystemExit):
        train.get_args(["--dataset", "imagenet"])


def test_synthetic_dataset_is_deterministic_and_shaped():
    ds = train.SyntheticImages(32, 3, 16, seed=3)
    x, y = ds[5]
    x2, y2 = train.SyntheticImages(32, 3, 16, seed=3)[5]
    assert x.shape == (3, 16, 16) and 0 <= y < 10 and y == y2 and torch.equal(x, x2)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,amp", [("rope-axial", "none"), ("polynomial", "none"), ("rope-mixed", "bf16")])
def test_training_run_writes_reference_format_log_and_checkpoint(tmp_path, mode, amp):
    from vit_rpe_rope_b200.models.vit import VisionTransformer
    argv = ["--dataset", "cifar10", "--pos_encoding", mode, "--epochs", "3", "--batch_size", "64", "--synthetic",
            "--synthetic_size", "512", "--log_dir", str(tmp_path / "logs"), "--ckpt_dir", str(tmp_path / "ck"), "--amp", amp]
    if amp == "bf16":  # head dim 64: the tcgen05 kernels
        argv += ["--embed_dim", "128", "--num_heads", "2", "--depth", "2"]
    log_file, ckpt = train.main(argv)
    rows = list(csv.reader(open(log_file)))
    assert rows[0] == ["epoch", "train_loss", "train_acc", "test_loss", "test_acc", "best_acc"]
    assert [int(r[0]) for r in rows[1:]] == [1, 2, 3]
    losses = [float(r[1]) for r in rows[1:]]
    assert losses[-1] < losses[0], losses            # it learns the synthetic classes
    assert float(rows[-1][5]) >= max(float(r[4]) for r in rows[1:]) - 1e-9
    assert os.path.basename(ckpt) == f"cifar10_{mode}_best.pth"
    kw = dict(img_size=32, patch_size=4, in_chans=3, num_classes=10, pos_encoding=mode)
    if amp == "bf16":
        kw.update(embed_dim=128, num_heads=2, depth=2)
    VisionTransformer(**kw).load_state_dict(torch.load(ckpt), strict=True)
