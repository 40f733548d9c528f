"""The reference's single-GPU inference path around the network, restated on stock PyTorch.
MEASUREMENT / TEST INFRASTRUCTURE (see __init__.py): the product never imports this.

What the reference does for one network on a CUDA device (file:line relative to
/root/reference/body_organ_analysis/_external/):

  nnunetv2/inference/predict_from_raw_data.py
    :60      torch.backends.cudnn.benchmark = True (set by the TotalSegmentator wrapper before predicting)
    :648     the whole prediction runs under torch.autocast("cuda")  -> fp16 convs / norms / logits
    :568-582 a producer THREAD clones one patch at a time to the device through a Queue(maxsize=2)
    :587-590 accumulators `predicted_logits` [C, X, Y, Z] and `n_predictions` [X, Y, Z] are torch.half on the device
    :603-614 per patch: network(x)[0]; prediction *= gaussian; predicted_logits[sl] += prediction;
             n_predictions[sl[1:]] += gaussian          (three elementwise launches per patch)
    :620-625 torch.div(predicted_logits, n_predictions, out=predicted_logits); isinf -> RuntimeError
    :494-500 fold ensemble: every fold's logits go `.to('cpu')`, are summed there and divided by the fold count
    :386     `.cpu().numpy()` of the C x V logits
  nnunetv2/inference/export_prediction.py:38 -> label_handling.py:178   numpy argmax(0) on the host
  totalsegmentator/nnunet.py:553-556          part merge: one masked write per class of every part model
"""
from __future__ import annotations

from queue import Queue
from threading import Thread

import numpy as np
import torch

from oracle.sliding_window import compute_gaussian, pad_nd_image, sliding_window_slicers


@torch.inference_mode()
def predict_sliding_window_return_logits(network: torch.nn.Module, data: torch.Tensor, patch, step: float, device,
                                         use_gaussian: bool = True, autocast: bool = True,
                                         patch_range=None, accumulator_dtype=torch.half) -> torch.Tensor:
    """data [1, x, y, z] fp32 (host) -> logits [C, x, y, z] `accumulator_dtype` ON THE DEVICE, as
    nnUNetPredictor.predict_sliding_window_return_logits returns them (predict_from_raw_data.py:634-680)."""
    device = torch.device(device)
    padded, unpad = pad_nd_image(data.numpy() if isinstance(data, torch.Tensor) else data, patch)
    slicers = sliding_window_slicers(padded.shape[1:], patch, step)
    if patch_range is not None:
        slicers = slicers[patch_range[0]:patch_range[1]]
    with torch.autocast(device.type, enabled=autocast):
        d = torch.from_numpy(np.ascontiguousarray(padded, dtype=np.float32)).to(device)
        queue: Queue = Queue(maxsize=2)

        def producer():
            for sl in slicers:
                queue.put((torch.clone(d[(slice(None), *sl)][None], memory_format=torch.contiguous_format).to(device), sl))
            queue.put("end")

        t = Thread(target=producer)
        t.start()
        n_heads = None
        predicted_logits = n_predictions = None
        g = torch.from_numpy(compute_gaussian(tuple(patch), 1.0 / 8, 10)).to(device) if use_gaussian else 1
        if use_gaussian and accumulator_dtype != torch.half:
            g = g.to(accumulator_dtype)
        while True:
            item = queue.get()
            if isinstance(item, str):
                queue.task_done()
                break
            workon, sl = item
            prediction = network(workon)[0]
            if predicted_logits is None:
                n_heads = prediction.shape[0]
                predicted_logits = torch.zeros((n_heads, *padded.shape[1:]), dtype=accumulator_dtype, device=device)
                n_predictions = torch.zeros(padded.shape[1:], dtype=accumulator_dtype, device=device)
            prediction = prediction.to(accumulator_dtype) if prediction.dtype != accumulator_dtype else prediction
            if use_gaussian:
                prediction *= g
            predicted_logits[(slice(None), *sl)] += prediction
            n_predictions[sl] += g
            queue.task_done()
        queue.join()
        t.join()
        torch.div(predicted_logits, n_predictions, out=predicted_logits)
        if patch_range is None and torch.any(torch.isinf(predicted_logits)):
            raise RuntimeError("Encountered inf in predicted array. Aborting...")
    return predicted_logits[unpad] if patch_range is None else predicted_logits


@torch.inference_mode()
def predict_logits_from_preprocessed_data(networks, data: torch.Tensor, patch, step: float, device,
                                          autocast: bool = True) -> torch.Tensor:
    """Fold ensemble on the HOST (predict_from_raw_data.py:471-504): [C, x, y, z] on the CPU."""
    prediction = None
    for net in networks:
        p = predict_sliding_window_return_logits(net, data, patch, step, device, autocast=autocast).to("cpu")
        prediction = p if prediction is None else prediction + p
    if len(networks) > 1:
        prediction /= len(networks)
    return prediction


def convert_logits_to_segmentation(logits_cpu: torch.Tensor) -> np.ndarray:
    """`.cpu().numpy()` + numpy argmax (predict_from_raw_data.py:386, label_handling.py:178, export_prediction.py:46)."""
    return logits_cpu.numpy().argmax(0).astype(np.uint8)


def merge_part(seg_combined: np.ndarray, seg: np.ndarray, lut) -> None:
    """totalsegmentator/nnunet.py:553-556: one masked write per class of the part model."""
    for jdx, gid in enumerate(lut):
        if jdx == 0:
            continue
        seg_combined[seg == jdx] = gid
