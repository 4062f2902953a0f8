"""GPU: the asynchronous precomputed-item feed (finetrainers_b200/data.py) under a SLOW consumer.  The consumer's stream
lags many items behind the host (as a CUDA-graph-replayed training step does): the device tensors the reader hands out must
stay intact until the consumer's queued work has read them (record_stream hand-over), while the loader thread keeps
recycling its pinned staging ring."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_precomputed_reader_async_feed_with_lagging_consumer(tmp_path):
    from finetrainers_b200.data import PrecomputedReader, PrecomputedOnceReader, save_item
    n = 24
    for i in range(n):
        save_item({"latents": torch.full((1, 64, 4, 16, 16), float(i)), "num_frames": 4,
                   "latents_mean": torch.zeros(1, 64), "latents_std": torch.ones(1, 64)}, i, tmp_path, "latent")
    dev = torch.device("cuda", 0)
    rd = PrecomputedReader(tmp_path, "latent", rank=0, world_size=1, device=dev, prefetch=3)
    static = torch.zeros(1, 64, 4, 16, 16, device=dev)
    sums = torch.zeros(n, device=dev)
    torch.cuda._sleep(200_000_000)          # ~0.1 s: the consumer stream starts far behind the host
    seen = 0
    for i, item in enumerate(rd):
        assert item["latents"].is_cuda and item["num_frames"] == 4
        static.copy_(item["latents"], non_blocking=True)   # what SFTTrainStep.micro_step does with the batch
        torch.cuda._sleep(2_000_000)                        # the "training step": the host runs ahead of it
        sums[i] = static.mean()
        del item
        seen += 1
    assert seen == n and rd.requires_data
    torch.cuda.synchronize()
    assert sums.tolist() == [float(i) for i in range(n)]
    # infinite variant: cycles over the rank's slice, never asks for more data
    once = PrecomputedOnceReader(tmp_path, "latent", rank=1, world_size=2, device=dev, prefetch=2)
    it = iter(once)
    got = [int(next(it)["latents"].flatten()[0].item()) for _ in range(15)]
    assert got == [12 + (k % 12) for k in range(15)] and not once.requires_data
    it.close()
