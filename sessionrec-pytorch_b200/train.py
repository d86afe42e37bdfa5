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


class TrainRunner:
    """`TrainRunner` (`src/utils/train.py:57-127`) on the fused step: same constructor and the same epoch loop - evaluate
    first, one pass over `train_loader`, `StepLR(step_size=3, gamma=0.1)` per epoch, evaluate, stop after `patience` epochs in
    which MRR and HR both fell - with `model.train_step(batch)` (zero_grad + forward + nll_loss + backward + Adam with the
    `fix_weight_decay` groups, one host call) in place of the five-line loop body (`:95-101`).  The running loss stays on the
    device and is read once per `log_interval` instead of `loss.item()` every batch.  `train_loader` yields what the
    reference's loaders yield, `([batch, ...], labels)`, or bare batches (`loader.BatchPrefetcher`, `loader.EpochBatches`)."""

    STEP_SIZE, GAMMA = 3, 0.1                 # `train.py:75`

    def __init__(self, dataset, model, train_loader, test_loader, device, lr=1e-3, weight_decay=0, patience=3):
        self.dataset, self.model = dataset, model
        self.train_loader, self.test_loader, self.device = train_loader, test_loader, device
        self.lr0, self.patience = lr, patience
        self.epoch = self.batch = 0
        model.configure_optimizer(lr=lr, weight_decay=weight_decay)

    def _lr(self, epochs_done):
        return self.lr0 * self.GAMMA ** (epochs_done // self.STEP_SIZE)

    def train(self, epochs, log_interval=100, log=print):
        import time
        max_mrr = max_hit = 0
        bad_counter = 0
        t = time.time()
        acc = None
        evaluate(self.model, self.test_loader, self.device)           # the reference evaluates once before training (`:91`)
        for _ in range(epochs):
            self.model.train()
            for batch in self.train_loader:
                if isinstance(batch, (tuple, list)):
                    batch = batch[0][0]
                loss = self.model.train_step(batch.to(self.device))
                acc = loss.detach().clone() if acc is None else acc + loss.detach()
                if self.batch > 0 and self.batch % log_interval == 0:
                    log(f'Batch {self.batch}: Loss = {float(acc) / log_interval:.4f}, Time Elapsed = {time.time() - t:.2f}s')
                    t = time.time()
                    acc = None
                self.batch += 1
            self.model.set_lr(self._lr(self.epoch + 1))           # scheduler.step() (`:111`)
            mrr, hit = evaluate(self.model, self.test_loader, self.device)
            log(f'Epoch {self.epoch}: MRR = {mrr * 100:.3f}%, Hit = {hit * 100:.3f}%')
            if mrr < max_mrr and hit < max_hit:
                bad_counter += 1
                if bad_counter == self.patience:
                    break
            else:
                bad_counter = 0
            max_mrr = max(max_mrr, mrr)
            max_hit = max(max_hit, hit)
            self.epoch += 1
        return max_mrr, max_hit
