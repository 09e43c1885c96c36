"""Deep-Compression style weight sharing (reference: pipeline/quant/kmeans.py, kmeans_helper.py).

The reference re-clusters every non-skipped nn.Linear weight in a forward-pre-hook: a [n_w, 255] distance
matrix per Lloyd iteration in PyTorch.  Here the same algorithm is one library call per layer
(sirenb200_kmeans_quantize: label / ordered cluster sums / centre update / codebook / predict kernels) run
as a "weight-load" transform right before the fused forward stages the weights.
"""
import torch
from torch import nn

from ... import _lib
from ... import engine as _engine


class KmeansQuant:
    def __init__(self, model, optim, bits=5, skip_ll=None):
        self.model, self.optim, self.bits = model, optim, bits
        self.skip_ll = list(skip_ll) if skip_ll is not None else ["layers.0.linear", "layers.7.linear"]
        self._hooks = []
        self._registered = False
        if hasattr(model, "_weight_transforms"):
            # fused Siren: nn.Linear.__call__ is bypassed, so the transform is attached to the model
            model._weight_transforms.append(self._requantize_all)
            model._post_backward.append(self._centroid_sgd_all)
            self._registered = True
        else:
            for name, module in model.named_modules():
                if name not in self.skip_ll and isinstance(module, nn.Linear):
                    self._hooks.append(module.register_forward_pre_hook(self.kmeans_modify_weight))

    @property
    def n_clusters(self):
        return 2 ** self.bits

    @property
    def learning_rate(self):
        return self.optim.defaults["lr"]

    def _targets(self):
        return [(n, m) for n, m in self.model.named_modules()
                if n not in self.skip_ll and isinstance(m, nn.Linear)]

    def find_centroids(self, module):
        """kmeans.py:110-150 on the device: returns (centroids, labels, new_weight)."""
        weight = module.weight.data
        _lib.require_cuda(weight, "weight")
        return _engine.kmeans_quantize(weight, self.bits)

    def kmeans_modify_weight(self, module, input=None):
        """kmeans.py:65-71."""
        centroids, labels, new_weight = self.find_centroids(module)
        module.labeled_weight = labels
        module.centroids = centroids
        module.weight.data = new_weight

    def _requantize_all(self, model):
        for _, module in self._targets():
            self.kmeans_modify_weight(module)

    def scalar_quantization(self, module):
        """kmeans.py:152-177 (the reference's module backward hook): one SGD step on the code book,
        centroids <- centroids - lr * scatter_add(labels, dL/dW).  Its effect lasts until the next forward
        re-clusters the layer."""
        dw = torch.zeros_like(module.centroids)
        dw.scatter_add_(0, module.labeled_weight.flatten(), module.weight.grad.flatten())
        module.centroids = module.centroids - self.learning_rate * dw

    def _centroid_sgd_all(self, model):
        for _, module in self._targets():
            if hasattr(module, "labeled_weight") and module.weight.grad is not None:
                self.scalar_quantization(module)

    def labels_to_weights(self, labeled_weight, centroids):
        return centroids[labeled_weight]

    def update_weights(self):
        """kmeans.py:73-98: freeze the code book: centroids / labeled_weight become (non-trainable)
        Parameters and the weight is rebuilt from them."""
        for h in self._hooks:
            h.remove()
        self._hooks = []
        if self._registered:
            self.model._weight_transforms.remove(self._requantize_all)
            self.model._post_backward.remove(self._centroid_sgd_all)
            self._registered = False
        for _, module in self._targets():
            centroids, labels = module.centroids, module.labeled_weight
            module.centroids = nn.Parameter(centroids, requires_grad=False)
            module.labeled_weight = nn.Parameter(labels, requires_grad=False)
            module.weight.data = self.labels_to_weights(labels, centroids)

    def remove_labels_centroids(self):
        for _, module in self._targets():
            del module.labeled_weight
            del module.centroids
