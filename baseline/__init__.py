"""MEASUREMENT / TEST INFRASTRUCTURE - never imported by the product (`body-and-organ-analysis_b200/`).

The reference's single-GPU path, restated on stock PyTorch (cuDNN under `torch.autocast`), so that the "x times the
reference's PyTorch/nnU-Net path" figure of BASELINE.json has a measured denominator on the same box:

  torch_unet.py       PlainConvUNet as a torch.nn module tree laid out like dynamic_network_architectures==0.4.3
                      (the class the reference instantiates at _external/nnunetv2/utilities/get_network_from_plans.py:9-43;
                      the package itself is not installable offline)
  reference_loop.py   the reference's sliding-window loop and export steps around it
                      (_external/nnunetv2/inference/predict_from_raw_data.py:471-504,560-680,
                       _external/nnunetv2/inference/export_prediction.py:14-71,
                       _external/totalsegmentator/nnunet.py:534-559)

Users: bench.py (`gpu_baseline` leg), tests/ (the reference-GPU-path numerics the parity bar is stated against).
`baseline/_ref/` (git-ignored) is where an installed copy of the reference would live; it cannot be installed here
(DESIGN.md).
"""
