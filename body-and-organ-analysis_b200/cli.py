"""Command line: the flags of `python -m body_organ_analysis` (body_organ_analysis/cli.py:37-278) that concern the
accelerated path.  `python -m boa_b200 --input-image X.nii.gz --models total+bca -d gpu -o OUT`."""
from __future__ import annotations

import argparse
import logging
import os
from pathlib import Path

from .config import AVAILABLE_MODELS, env_bool, resolve_device, resolve_models


def _validate_models(spec: str) -> str:
    try:
        resolve_models(spec, strict=True)
    except ValueError as e:
        raise argparse.ArgumentTypeError(str(e)) from e
    return spec


def get_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="boa_b200")
    p.add_argument("-i", "--input-image", default="/dicoms", type=Path, help="Path to the NIfTI file")
    p.add_argument("-o", "--output-dir", default="/workspace", type=Path)
    p.add_argument("-m", "--models", required=True, type=_validate_models, metavar="MODEL[+MODEL...]",
                   help=f"Plus-separated list of models, e.g. 'total+bca'. Available: {', '.join(sorted(AVAILABLE_MODELS))}.")
    p.add_argument("-d", "--device", default=None, type=str, help="'gpu', 'cuda', 'gpu:<id>' (there is no CPU path)")
    p.add_argument("--fast-bca", default=None, action="store_true")
    p.add_argument("--fast-total", default=None, action="store_true")
    p.add_argument("--bca-no-pdf", default=None, action="store_true", help="accepted; the PDF report is never produced")
    p.add_argument("--bca-median-filtering", default=False, action="store_true",
                   help="in-plane 3x3 median of the CT before the tissue HU thresholds")
    p.add_argument("--cnr-adjustment", default=None, action="store_true")
    p.add_argument("--skip-contrast-information", default=None, action="store_true")
    p.add_argument("--force-recompute", default=False, action="store_true")
    p.add_argument("--weights", default=None, type=str, help="weights directory (default: $TOTALSEG_WEIGHTS_PATH)")
    p.add_argument("-v", "--verbose", default=False, action="store_true")
    return p


def run(argv=None) -> None:
    args = get_parser().parse_args(argv)
    logging.basicConfig(level=logging.DEBUG if args.verbose else logging.INFO)
    device = resolve_device(args.device)
    models = resolve_models(args.models, strict=True)
    fast_bca = args.fast_bca if args.fast_bca is not None else env_bool("FAST_BCA")
    fast_total = args.fast_total if args.fast_total is not None else env_bool("FAST_TOTAL")
    from .commands import analyze_ct
    out, stats = analyze_ct(args.input_image, args.output_dir, models=models, fast_bca=fast_bca, fast_total=fast_total,
                            bca_median_filtering=bool(args.bca_median_filtering),
                            cnr_adjustment=bool(args.cnr_adjustment), device=device,
                            recompute=args.force_recompute, weights_root=args.weights or os.environ.get("TOTALSEG_WEIGHTS_PATH"))
    logging.getLogger(__name__).info("results in %s: %s", out, stats)
