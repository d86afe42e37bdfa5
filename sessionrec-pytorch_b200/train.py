"""Drop-in for the two helpers of `src/utils/train.py` that touch the hot path: `prepare_batch` (`:26-32`) and
`evaluate` (`:36-55`, MRR@cutoff / HR@cutoff).  `evaluate` uses the fused scoring + top-k path (model.topk) when the
model offers it, else the reference arithmetic on the (B, V) log-probabilities."""
import torch


def prepare_batch(batch, device):
    inputs, labels = batch
    return [x.to(device) for x in inputs], labels.to(device)


def evaluate(model, data_loader, device, cutoff=20):
    model.eval()
    mrr = hit = num_samples = 0
    with torch.no_grad():
        for batch in data_loader:
            inputs, labels = prepare_batch(batch, device)
            if hasattr(model, 'topk'):
                topk = model.topk(*inputs, k=cutoff)
            else:
                topk = model(*inputs).topk(k=cutoff)[1]
            num_samples += topk.size(0)
            hit_ranks = torch.where(topk == labels.unsqueeze(-1))[1] + 1
            hit += hit_ranks.numel()
            mrr += hit_ranks.float().reciprocal().sum().item()
    return mrr / num_samples, hit / num_samples
