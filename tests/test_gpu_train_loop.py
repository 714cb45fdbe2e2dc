"""GPU: the drop-in model inside the reference's training-loop contract (pyrutils/torch/train_utils.py:118-165 restated in
2g-gcn_b200/train_loop.py): dataset tuples -> fetcher -> feeder -> criterion list -> backward -> clip -> Adam, several
batches; the summed loss must go down when the same small dataset is revisited."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_training_loop_reduces_the_loss(orc, synth, pkg):
    shape = synth.SHAPES['mphoi']
    D, B, T, n_videos = 32, 4, 12, 8
    torch.manual_seed(3)
    model = pkg.TGGCN(**synth.model_kwargs(shape, hidden_size=D, stage=2)).cuda()
    batch = synth.make_batch(shape, n_videos, T, seed=21)
    targets = synth.target_list(shape, synth.make_targets(shape, batch['lengths'], T, seed=22))
    # tuple layout of the reference's TensorDataset (vhoi/data_loading.py:517-519): 8 inputs then the targets
    zeros = torch.zeros(n_videos, 1)
    tensors = [batch['x_human'], batch['x_objects'], batch['objects_mask'], zeros, zeros, zeros, zeros, batch['steps_per_example']]
    dataset = torch.utils.data.TensorDataset(*tensors, *targets)
    loader = torch.utils.data.DataLoader(dataset, batch_size=B, shuffle=False)

    def fetch(ds, device):           # stand-in for gcn_fetcher (vhoi/data_loading.py:1282-1315)
        ds = [t.to(device) for t in ds]
        return ds[:8], ds[8:]

    def feed(m, data):               # stand-in for gcn_forward (vhoi/data_loading.py:1233-1279), stage-2 settings
        return m(x_human=data[0], x_objects=data[1], objects_mask=data[2], human_segmentation=None,
                 steps_per_example=data[7], inspect_model=False)

    def criterion(output, target, reduction='mean'):
        return orc.multi_task_loss(output, target, 'mphoi', 2)

    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    reducer = pkg.dp.GradientAllReduce(model)          # world size 1: rebinding only
    torch.manual_seed(5)                               # the model draws its Gumbel noise from the global CPU generator
    epochs = []
    for _ in range(6):
        hist = pkg.train_loop.train_single_epoch(model, loader, opt, criterion, 'cuda', ['hb', 'hs', 'fr', 'fp', 'sr', 'sp'],
                                                 clip_gradient_at=5.0, fetch_model_data=fetch, feed_model_data=feed,
                                                 reducer=reducer, verbose=False)
        epochs.append(float(torch.stack(hist).sum()))
    model.check_persistent_kernels()
    assert all(torch.isfinite(torch.tensor(epochs)))
    assert epochs[-1] < 0.92 * epochs[0] and epochs[-1] < epochs[2] < epochs[0], epochs
