"""gpytorch.lazy.LazyEvaluatedKernelTensor, dense edition.

``Kernel.__call__`` returns this object; slicing its last two dimensions slices the INPUTS and the
kernel is evaluated on the sub-blocks only when densified.  That is what makes
``VariationalStrategy.forward`` evaluate K(Z,Z), K(Z,X) and K(X,X) as three separate kernel calls
(each with its own ``x1`` centring inside ``sq_dist`` / ``MaternKernel.forward``).
"""
import torch


class LazyEvaluatedKernelTensor:
    def __init__(self, x1, x2, kernel, last_dim_is_batch=False, **params):
        self.x1, self.x2, self.kernel, self.params = x1, x2, kernel, params
        self.last_dim_is_batch = last_dim_is_batch

    @property
    def shape(self):
        return torch.Size([*torch.broadcast_shapes(self.x1.shape[:-2], self.x2.shape[:-2]), self.x1.shape[-2], self.x2.shape[-2]])

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    @property
    def dtype(self):
        return self.x1.dtype

    @property
    def device(self):
        return self.x1.device

    def __getitem__(self, index):
        if not isinstance(index, tuple):
            index = (index,)
        if len(index) != 3 or index[0] is not Ellipsis or not all(isinstance(i, slice) for i in index[1:]):
            raise NotImplementedError("shim: only [..., rows, cols] slicing is supported")
        row, col = index[1], index[2]
        return LazyEvaluatedKernelTensor(self.x1[..., row, :], self.x2[..., col, :], self.kernel,
                                         last_dim_is_batch=self.last_dim_is_batch, **self.params)

    def evaluate_kernel(self):
        """LazyEvaluatedKernelTensor.evaluate_kernel: kernel(x1, x2) with lazy evaluation switched off."""
        return self.kernel._evaluate(self.x1, self.x2, **self.params)

    def to_dense(self):
        return self.evaluate_kernel()

    def add_jitter(self, jitter_val=1e-3):
        """LinearOperator.add_jitter: + jitter * I (evaluates the kernel)."""
        dense = self.evaluate_kernel()
        n = dense.shape[-1]
        return dense + torch.eye(n, dtype=dense.dtype, device=dense.device) * jitter_val
