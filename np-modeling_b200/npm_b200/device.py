"""DeviceArray — the ndarray stand-in that flows between layers.

The reference passes NumPy arrays between `Layer.forward/backward` calls
(layers/layer.py:19-25).  Here the same objects flow, but the bytes live in B200
HBM: a DeviceArray is a thin owner of a contiguous fp32 CUDA buffer (allocated
through torch's caching allocator — torch is plumbing only) exposing the small
ndarray surface the reference's glue code uses (`shape`, `size`, `reshape`,
`+=`, indexing on the leading axis, `np.asarray(x)`).

Every arithmetic operation on it is one of this repo's CUDA kernels, reached
through the C-ABI; nothing here computes on the host.
"""
import os

import numpy as np
import torch

from ._lib import C

_NO_OPT_PLANES = bool(os.environ.get('NPM_NO_OPT_PLANES'))     # A/B: the optimizer does not maintain weight planes


def stream():
    """cudaStream_t (as int) all kernels are enqueued on: torch's current stream."""
    return torch.cuda.current_stream().cuda_stream


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('np-modeling_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


class DeviceArray:
    # `colsum`: optional DeviceArray [shape[-1]] holding the sums over all leading axes, attached by a kernel that had
    # the values in registers anyway (the fused LayerNorm backward); consumers that need exactly that reduction — the
    # bias gradient `np.sum(dy, axis=0)` of mlp.py:34 / attentions.py:129 — use it instead of re-reading the array.
    # The attachment lives in a cell shared by every reshape() alias / leading-axis view of the same storage, so an
    # in-place write through ANY of them invalidates it for all of them.
    __slots__ = ('t', '_cs')
    __array_priority__ = 1000

    def __init__(self, t: torch.Tensor, colsum=None, _cell=None):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), 'DeviceArray wraps contiguous fp32 CUDA'
        self.t = t
        # [colsum DeviceArray, numel of the array it sums, weight planes [buffer, base pointer, base numel, fresh, in sync] or None]
        self._cs = _cell if _cell is not None else [None, 0, None]
        if colsum is not None:
            self.colsum = colsum

    @property
    def colsum(self):
        cs, numel = self._cs[0], self._cs[1]
        if cs is None or numel != self.t.numel() or self.t.dim() < 1 or cs.size != self.t.shape[-1]:
            return None
        return cs

    @colsum.setter
    def colsum(self, value):
        self._cs[0] = value
        self._cs[1] = self.t.numel() if value is not None else 0

    def touched(self):
        """Call after writing the storage behind this array other than through its own methods (a raw write through
        `.t`): drops everything derived from the old contents (column sums, the split-bf16 image of a weight)."""
        self._cs[0], self._cs[1] = None, 0
        if self._cs[2] is not None:
            self._cs[2][3] = self._cs[2][4] = False

    # ---- ndarray-like surface -------------------------------------------------
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def ndim(self):
        return self.t.dim()

    @property
    def size(self):
        return self.t.numel()

    @property
    def dtype(self):
        return np.dtype(np.float32)

    @property
    def ptr(self):
        return self.t.data_ptr()

    def __len__(self):
        return self.t.shape[0]

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __repr__(self):
        return f'DeviceArray(shape={self.shape}, device={self.t.device})'

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return DeviceArray(self.t.view(*shape), _cell=self._cs)

    def copy(self):
        return DeviceArray(self.t.clone())

    def __deepcopy__(self, memo):
        return self.copy()

    def __getitem__(self, idx):
        v = self.t[idx]
        if not v.is_contiguous():
            raise IndexError('DeviceArray only supports contiguous (leading-axis) views')
        return DeviceArray(v, _cell=self._cs)     # a write through the view invalidates the parent's column sums

    def __iadd__(self, other):
        other = asdevice(other)
        assert other.shape == self.shape, f'{other.shape} vs {self.shape}'
        C.npm_add_inplace(self.ptr, other.ptr, self.size, stream())
        self.touched()
        return self

    def __add__(self, other):
        other = asdevice(other)
        assert other.shape == self.shape, f'{other.shape} vs {self.shape}'
        out = empty(self.shape)
        C.npm_add3(self.ptr, other.ptr, None, out.ptr, self.size, stream())
        return out

    # The rest of the arithmetic a user-defined Optimizer written in the reference's style needs
    # (`variable -= lr * gradient`, optimizer.py:32): each is one or two of this library's elementwise kernels.
    def _scaled(self, s: float):
        out = self.copy()
        C.npm_scale(out.ptr, float(s), out.size, stream())
        return out

    def __mul__(self, other):
        if isinstance(other, (int, float, np.floating, np.integer)):
            return self._scaled(float(other))
        return NotImplemented

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, (int, float, np.floating, np.integer)):
            return self._scaled(1.0 / float(other))
        return NotImplemented

    def __neg__(self):
        return self._scaled(-1.0)

    def __sub__(self, other):
        return self + (-asdevice(other))

    def __isub__(self, other):
        self += -asdevice(other)
        return self

    def __imul__(self, other):
        if not isinstance(other, (int, float, np.floating, np.integer)):
            return NotImplemented
        C.npm_scale(self.ptr, float(other), self.size, stream())
        self.touched()
        return self

    def item(self):
        return float(self.t.item())

    def __float__(self):
        return self.item()

    def copy_from(self, src):
        """In-place overwrite from a host array or another DeviceArray (keeps the buffer identity)."""
        if isinstance(src, DeviceArray):
            self.t.copy_(src.t.view(self.t.shape))
        else:
            host = np.ascontiguousarray(np.asarray(src), dtype=np.float32).reshape(self.shape)
            self.t.copy_(torch.from_numpy(host), non_blocking=False)
        self.touched()
        return self


class PlanesArray:
    """An activation that exists ONLY as split-bf16 planes — bf16 hi = bf16_rn(x) and mid = bf16_rn(x - hi), two
    contiguous [rows, cols] bf16 matrices in one buffer — because the kernel that produced it wrote it that way (the first
    FFN GEMM's epilogue, the ReLU backward) and the GEMMs that consume it land it by TMA as their operand image without
    converting (gemm_bx.cu A_PRE / B_PRE).  Only the split-bf16 ('bf16x3') mode makes one, and only between layers of this
    package that asked for it (`_planes_ok`); everything else goes through `to_fp32()` (exact: 16 significant bits).
    There is deliberately no `.ptr`: a consumer that is not planes-aware fails loudly instead of reading garbage."""
    __slots__ = ('buf', '_shape')
    __array_priority__ = 1000

    def __init__(self, buf: torch.Tensor, shape):
        self.buf = buf
        self._shape = tuple(int(v) for v in shape)
        assert buf.dtype == torch.uint8 and buf.numel() >= 4 * self.size

    shape = property(lambda self: self._shape)
    ndim = property(lambda self: len(self._shape))
    size = property(lambda self: int(np.prod(self._shape)) if self._shape else 1)
    dtype = property(lambda self: np.dtype(np.float32))
    hi_ptr = property(lambda self: self.buf.data_ptr())      # the bf16 hi plane; the mid plane starts `size` elements later
    colsum = None

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = tuple(int(v) for v in shape)
        if -1 in shape:
            known = int(np.prod([v for v in shape if v != -1]))
            shape = tuple(self.size // known if v == -1 else v for v in shape)
        assert int(np.prod(shape)) == self.size, f'cannot reshape {self._shape} to {shape}'
        return PlanesArray(self.buf, shape)

    def to_fp32(self) -> 'DeviceArray':
        out = empty(self._shape)
        C.npm_planes_join(self.hi_ptr, self.size, out.ptr, self.size, stream())
        return out

    def numpy(self):
        return self.to_fp32().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __len__(self):
        return self._shape[0]

    def __repr__(self):
        return f'PlanesArray(shape={self.shape}, split-bf16)'


def asdevice(x) -> DeviceArray:
    """np.ndarray / scalar sequence / torch tensor / DeviceArray → DeviceArray (fp32, on the current GPU).  A PlanesArray
    passes through (layers that take one check for it; the others call `.to_fp32()` via `asfp32`)."""
    if isinstance(x, (DeviceArray, PlanesArray)):
        return x
    if isinstance(x, torch.Tensor):
        return DeviceArray(x.detach().to(device=_device(), dtype=torch.float32).contiguous())
    host = np.ascontiguousarray(np.asarray(x), dtype=np.float32)
    return DeviceArray(torch.from_numpy(host).to(_device(), non_blocking=False))


def asfp32(x) -> DeviceArray:
    """asdevice for consumers that need the fp32 bytes: a PlanesArray is joined (one elementwise kernel)."""
    x = asdevice(x)
    return x.to_fp32() if isinstance(x, PlanesArray) else x


def from_pinned(host_pinned: torch.Tensor) -> DeviceArray:
    """Asynchronous H2D copy from a pinned host tensor on the current stream."""
    return DeviceArray(host_pinned.to(_device(), non_blocking=True))


def empty(shape) -> DeviceArray:
    if isinstance(shape, int):
        shape = (shape,)
    return DeviceArray(torch.empty(tuple(int(s) for s in shape), dtype=torch.float32, device=_device()))


def zeros(shape) -> DeviceArray:
    """Zero-filled by this library's own fill kernel (torch is the allocator only)."""
    out = empty(shape)
    if out.size:
        C.npm_fill(out.ptr, 0.0, out.size, stream())
    return out


def split_weight(w: 'DeviceArray'):
    """bf16 hi / mid planes of a weight for the split-bf16 ('bf16x3') contraction mode (npm_weight_split), or None in
    the other modes.  A layer asks once per forward call; its forward GEMM and its dX GEMM both take them as their B
    operand without converting it again (gemm_bx.cu, B_PRE).

    The buffer is kept on the array (shared by its views).  The fused optimizers rewrite it inside their update kernel
    (npm_tensor_entry.planes) and mark it fresh; a fresh image is handed out ONCE — the first forward after an update
    needs no split pass over the weights — and every other call splits again, so nothing that writes the weight in any
    other way in between (a raw write through `.t`, a user optimizer) can leave a stale image behind."""
    from ._lib import PREC_BF16X3, load
    if load().npm_get_precision() != PREC_BF16X3 or w.size % 4 or w.ndim < 2:
        return None
    rows, cols = w.shape[0], w.size // w.shape[0]
    wp = w._cs[2]
    if wp is not None and wp[1] == w.ptr and wp[2] == w.size:
        if wp[3] and not _NO_OPT_PLANES:
            wp[3] = False
            return wp[0]
        planes = wp[0]
    else:
        planes = torch.empty(int(C.npm_weight_split_bytes(rows, cols)), dtype=torch.uint8, device=_device())
        wp = w._cs[2] = [planes, w.ptr, w.size, False, False]
    C.npm_weight_split(w.ptr, planes.data_ptr(), rows, cols, stream())
    wp[4] = True          # the image matches the weight as of now
    return planes


def weight_planes_of(v: 'DeviceArray'):
    """(hi-plane pointer, plane stride in elements) of the split-bf16 image kept for parameter `v` (a whole array or a
    view into a packed block), or None: what the fused optimizers put into npm_tensor_entry.planes."""
    wp = v._cs[2]
    if wp is None or _NO_OPT_PLANES:
        return None
    buf, base, numel, _, synced = wp
    if not synced:
        return None           # the weight was written behind the image's back since the last split: the next forward splits again
    off = v.ptr - base
    if off < 0 or off + 4 * v.size > 4 * numel or off % 16 or v.size % 4 or v.ptr % 16:
        return None
    return buf.data_ptr() + off // 2, numel


def mark_weight_planes_fresh(v: 'DeviceArray'):
    """The update kernel has rewritten `v`'s part of the image (it was given weight_planes_of(v))."""
    if v._cs[2] is not None and v._cs[2][4]:
        v._cs[2][3] = True


def workspace(nbytes: int) -> torch.Tensor:
    """Scratch bytes for a *_workspace() query (caller-owned, per the C-ABI contract)."""
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=_device())


def param(x):
    """Normalise a parameter attribute: users may assign NumPy arrays (as the reference's tests do,
    layers/utils.py:41-101); they are moved to the device on first use."""
    return x if isinstance(x, DeviceArray) else asdevice(x)


class DeviceScalar:
    """A loss value still on the device.  Behaves like the Python float the reference returns
    (loss.py:25,36) but only synchronises with the GPU when somebody looks at it."""
    __slots__ = ('_t', '_v')
    __array_priority__ = 1000

    def __init__(self, arr: DeviceArray):
        self._t = arr.t
        self._v = None

    def __float__(self):
        if self._v is None:
            self._v = float(self._t.item())
        return self._v

    def item(self):
        return float(self)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(float(self), dtype=dtype or np.float32)

    def __repr__(self):
        return repr(float(self))

    __str__ = __repr__

    def __format__(self, spec):
        return format(float(self), spec)

    def __add__(self, o): return float(self) + float(o)
    __radd__ = __add__
    def __sub__(self, o): return float(self) - float(o)
    def __rsub__(self, o): return float(o) - float(self)
    def __mul__(self, o): return float(self) * float(o)
    __rmul__ = __mul__
    def __truediv__(self, o): return float(self) / float(o)
    def __lt__(self, o): return float(self) < float(o)
    def __le__(self, o): return float(self) <= float(o)
    def __gt__(self, o): return float(self) > float(o)
    def __ge__(self, o): return float(self) >= float(o)
    def __eq__(self, o): return float(self) == float(o)
    def __hash__(self): return hash(float(self))
