"""CPU: host logic of dfvod_b200.data_parallel.GradientAllReducer that needs no process group (world size 1)."""
import pytest
import torch

from dfvod_b200 import data_parallel


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))


def test_buckets_are_sent_in_index_order_and_gradients_survive():
    model = _model()
    reducer = data_parallel.GradientAllReducer(model.parameters(), bucket_bytes=256)    # several small buckets
    assert len(reducer.buckets) > 1
    order = []
    launch = reducer._launch
    reducer._launch = lambda b: (order.append(b), launch(b))[1]
    x = torch.randn(5, 8)
    model(x).square().mean().backward()
    want = [p.grad.clone() for p in model.parameters()]
    reducer.finish()
    assert order == sorted(order) == list(range(len(reducer.buckets)))        # same order on every rank
    for p, w in zip(model.parameters(), want):
        assert torch.equal(p.grad, w)
    reducer.remove()


def test_unused_parameters_are_flushed_by_finish_in_order():
    model = _model()
    extra = torch.nn.Linear(3, 3)                                  # never touched by the loss
    params = list(model.parameters()) + list(extra.parameters())
    reducer = data_parallel.GradientAllReducer(params, bucket_bytes=256)
    order = []
    launch = reducer._launch
    reducer._launch = lambda b: (order.append(b), launch(b))[1]
    model(torch.randn(2, 8)).sum().backward()
    reducer.finish()
    assert order == list(range(len(reducer.buckets)))
    assert all(p.grad is not None and not bool(p.grad.any()) for p in extra.parameters())
    reducer.remove()


def test_second_backward_before_finish_is_an_error():
    model = _model()
    reducer = data_parallel.GradientAllReducer(model.parameters())
    x = torch.randn(3, 8)
    model(x).sum().backward()
    with pytest.raises(RuntimeError, match="finish"):
        model(x).sum().backward()
    reducer.remove()
