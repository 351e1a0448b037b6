"""``python -m tbmodels_b200 eigenvals`` -- the reference's ``tbmodels eigenvals`` command (src/tbmodels/_cli.py:227-262)
with the same options, defaults and messages, evaluated on the GPU and without the h5py / bands_inspect dependencies."""
from __future__ import annotations

import argparse
import sys


def _eigenvals(args) -> int:
    import numpy as np

    from . import Evaluator, io

    if args.verbose:
        print(f"Reading initial model from file '{args.input}' ...")
    model = io.load_model(args.input)
    if args.verbose:
        print(f"Reading kpoints from file '{args.kpoints}' ...")
    kpts = io.load_kpoints(args.kpoints)
    if args.verbose:
        print("Calculating energy eigenvalues ...")
    ev = Evaluator(model, device=args.device)
    try:
        eigenvalues = ev.eigenval_array(kpts) if len(kpts) else np.zeros((0, model.size))
    finally:
        ev.close()
    if args.verbose:
        print(f"Writing kpoints and energy eigenvalues to file '{args.output}' ...")
    io.save_eigenvals(args.output, kpts, eigenvalues)
    if args.verbose:
        print("Done!")
    return 0


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="tbmodels_b200", description="GPU evaluator for TBmodels tight-binding models.")
    sub = ap.add_subparsers(dest="command", required=True)
    ev = sub.add_parser("eigenvals", help="Calculate energy eigenvalues.",
                        description="Calculate the energy eigenvalues for a given set of k-points (in reduced "
                                    "coordinates). The input and output is given in an HDF5 file.")
    ev.add_argument("-i", "--input", default="model.hdf5", help="File containing the input model (in HDF5 format).")
    ev.add_argument("-k", "--kpoints", default="kpoints.hdf5",
                    help="File containing the k-points for which the eigenvalues are evaluated.")
    ev.add_argument("-o", "--output", default="eigenvals.hdf5", help="Output file for the energy eigenvalues.")
    ev.add_argument("-v", "--verbose", action="store_true", help="Enable verbose output.")
    ev.add_argument("--device", type=int, default=None, help="CUDA device index (default: current device).")
    ev.set_defaults(func=_eigenvals)
    args = ap.parse_args(argv)
    return args.func(args)


if __name__ == "__main__":
    sys.exit(main())
