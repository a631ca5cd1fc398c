"""core23-style tensor / buffer runtime (lazy tensors, buffer channels, unitary buffers, containers)."""
import pytest
import torch

from hugectr_b200 import core23 as c


@pytest.fixture(autouse=True)
def _fresh():
    c.ReleaseBuffers()
    yield
    c.ReleaseBuffers()


def test_lazy_allocation_and_unitary_channel():
    dev = c.Device()
    wp = c.TensorParams(device=dev, buffer_params=c.BufferParams(c.BufferChannel("Weight")))
    a = c.Tensor((3, 5), torch.float32, wp)
    b = c.Tensor((7,), torch.float32, wp)
    other = c.Tensor((4,), torch.float32, c.TensorParams(device=dev, buffer_params=c.BufferParams(c.BufferChannel("Wgrad"))))
    assert a._data is None and b._data is None                 # nothing allocated yet
    assert a.num_elements() == 15 and a.num_bytes() == 60 and a.dims() == 2 and a.size(1) == 5
    a.data().fill_(1.0)                                         # first access allocates the whole channel
    assert b._data is not None and other._data is None
    buf = c.GetBuffer(wp.buffer_params, dev)
    flat = buf.decay(torch.float32)
    assert flat.numel() * 4 >= 60 + 28
    assert (b.data().data_ptr() - a.data().data_ptr()) % c.ALIGN == 0    # 256-byte aligned carving
    b.data().fill_(2.0)
    assert float(flat.sum()) == 15 * 1.0 + 7 * 2.0              # one flat array over both tensors
    flat.zero_()
    assert float(a.data().abs().sum()) == 0 and float(b.data().abs().sum()) == 0
    with pytest.raises(RuntimeError):
        buf.subscribe(c.Tensor.bind(torch.zeros(2)))            # channel already allocated
    late = c.Tensor((2,), torch.float32, wp)                    # a new generation of the channel
    assert late.data().numel() == 2
    assert c.AllocateBuffers(dev) and other._data is not None


def test_container_flat_view_bind_reshape_and_primitives():
    p = c.TensorParams(device=c.Device(), data_type=torch.float32,
                       buffer_params=c.BufferParams(c.GetRandomBufferChannel()))
    ts = [c.Tensor((4, 4), torch.float32, p), c.Tensor((10,), torch.float32, p), c.Tensor((1, 3), torch.float32, p)]
    cont = c.TensorContainer(ts, shape=(3,))
    flat = cont.flatten()
    flat.fill_(3.0)
    assert all(float(t.data().min()) == 3.0 for t in ts) and len(cont) == 3
    c.zeros_sync(ts[1])
    assert float(flat.sum()) == 3.0 * (16 + 3) + 3.0 * (flat.numel() - 16 - 10 - 3)   # padding counted
    r = ts[0].reshape((2, 8))
    assert not r.own_data() and r.data().data_ptr() == ts[0].data().data_ptr()
    with pytest.raises(ValueError):
        ts[0].reshape((5,))
    ext = torch.arange(6.0)
    bt = c.Tensor.bind(ext, shape=(2, 3))
    assert not bt.own_data() and bt.shape() == (2, 3)
    dst = c.Tensor((2, 3), torch.float32, p.with_(buffer_params=c.BufferParams(c.GetRandomBufferChannel())))
    c.copy_sync(dst, bt)
    assert torch.equal(dst.data(), ext.view(2, 3))
    half = c.Tensor((2, 3), torch.bfloat16, p.with_(data_type=torch.bfloat16,
                                                    buffer_params=c.BufferParams(c.GetRandomBufferChannel())))
    c.convert_async(half, dst)
    assert half.data().dtype == torch.bfloat16 and float(half.data().float().sum()) == 15.0
    g = torch.Generator().manual_seed(1)
    c.uniform_async(dst, -1, 1, g)
    assert float(dst.data().abs().max()) <= 1
    with pytest.raises(ValueError):
        c.TensorContainer([ts[0], half])
    # non-unitary buffers: one allocation per tensor, no flat view
    q = c.TensorParams(device=c.Device(), buffer_params=c.BufferParams(c.GetRandomBufferChannel(), unitary=False))
    x, y = c.Tensor((3,), torch.float32, q), c.Tensor((3,), torch.float32, q)
    x.data()
    with pytest.raises(RuntimeError):
        c.TensorContainer([x, y]).flatten()
