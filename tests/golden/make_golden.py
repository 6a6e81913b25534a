#!/usr/bin/env python3
"""Collects the golden vectors the reference ships for the ABneutral hot path.

Run in the build container (where /root/reference is mounted read-only):

    python tests/golden/make_golden.py

The files are DATA fixtures of alphabeta-rs (GPL-3.0, https://github.com/constantingoeldel/alphabeta-rs),
copied byte for byte so the tests can run where /root/reference does not exist (the GPU box).
No reference source code is copied.  MANIFEST.json records the sha256 of every fixture.

What pins what:
  pedigree.txt              Problem::default() input of the cost KAT   src/structs.rs:172-189,233
  divergence.txt            dt1t2 of the R original                     src/divergence.rs:139-161
  nodelist.txt edgelist.txt methylome/G*.txt -> pedigree_generated.txt  src/pedigree.rs:345-358
  desired_output/*          complete run of the R original (13 samples)  src/pedigree.rs:360-379 (disabled test)
  annotation.bed + output_metaplot/distribution*.txt  metaprofile example  README / src/cli/metaprofile.rs
"""
import hashlib
import json
import os
import shutil

REF = "/root/reference/data"
HERE = os.path.dirname(os.path.abspath(__file__))

FILES = [
    "pedigree.txt",
    "divergence.txt",
    "pedigree_generated.txt",
    "nodelist.txt",
    "edgelist.txt",
    "annotation.bed",
    "methylome/G0.txt",
    "methylome/G1_2.txt",
    "methylome/G4_2.txt",
    "methylome/G4_8.txt",
    "output_metaplot/distributions.txt",
    "output_metaplot/distribution_G0.txt",
    "output_metaplot/steady_state_methylation.txt",
    "desired_output/nodelist.fn",
    "desired_output/edgelist.fn",
    "desired_output/pedigree-pdata_epimutation_rate_estimation_window_gene_0.txt",
    "desired_output/p0uu_in_epimutation_rate_estimation_window_gene_0.txt",
    "desired_output/ABneutral_estimatats_epimutation_rate_estimation_window_gene_0.txt",
] + [
    f"desired_output/methylome_Col0_{g}_All.txt"
    for g in ["G0", "G1_L2", "G1_L8", "G2_L2", "G2_L8", "G4_L2", "G4_L8", "G5_L2", "G5_L8", "G8_L2", "G8_L8",
              "G11_L2", "G11_L8"]
]


def main():
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(HERE, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    # scalar known-answer values quoted from the reference's own tests
    manifest["_kat"] = {
        "cost_default_problem": "0.0006700888539608879  # src/structs.rs:233 (assert_eq!, exact)",
        "model_default": [0.0001179555, 0.0001180614, 0.03693534, 0.003023981],  # src/structs.rs:66-75
        "problem_default": {"p_uu": 0.75, "p_mm": 0.25, "eqp": 0.5, "eqp_weight": 0.7},  # src/structs.rs:172-189
        "same_as_r_params": {"p_mm": 0.25, "p_uu": 0.75, "alpha": 3.974271e-09, "beta": 1.519045e-07,
                             "weight": 0.06892953},  # src/divergence.rs:141-149
    }
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print(f"copied {len(FILES)} fixtures")


if __name__ == "__main__":
    main()
