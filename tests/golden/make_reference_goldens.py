"""Extracts the reference's own published example outputs into tests/golden/*.json.

Run in the build container (needs /root/reference); the JSON files it writes are committed so
that nothing at test time reads /root/reference.  Sources:
  * examples/README.md:7-12            (mvn_example stdout)
  * examples/multivariate_normal/mvn_example.ipynb   cell 4 stream output
  * examples/gaussian_mixture_model/gmm_example.ipynb cell 4 stream output
Model / optimizer parameters are the literals of examples/*/..._example.cpp.
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def notebook_stream(path):
    nb = json.load(open(path))
    for cell in nb["cells"]:
        for out in cell.get("outputs", []):
            if out.get("output_type") == "stream":
                return "".join(out["text"])
    raise RuntimeError("no stream output in " + path)


def parse_positions(text):
    """'Initial particle positions:\n [[..]..]\nFinal particle positions:\n [[..]]' -> two n x 2 lists."""
    blocks = re.split(r"(?:Initial|Final) particle positions:", text)[1:]
    res = []
    for b in blocks:
        nums = [float(t) for t in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?", b)]
        res.append([nums[i:i + 2] for i in range(0, len(nums), 2)])
    return res


def parse_readme(path):
    lines = open(path).read().splitlines()
    i0 = lines.index("Initial particle coordinates")
    i1 = lines.index("Final particle coordinates")
    rows = lambda k: [[float(t) for t in lines[k + r].split()] for r in (1, 2)]
    tr = lambda m: [[m[0][j], m[1][j]] for j in range(len(m[0]))]
    return tr(rows(i0)), tr(rows(i1))


mvn_init, mvn_final = parse_positions(notebook_stream(os.path.join(REF, "examples/multivariate_normal/mvn_example.ipynb")))
rd_init, rd_final = parse_readme(os.path.join(REF, "examples/README.md"))
assert rd_init == mvn_init and rd_final == mvn_final, "README and notebook disagree"
gmm_init, gmm_final = parse_positions(notebook_stream(os.path.join(REF, "examples/gaussian_mixture_model/gmm_example.ipynb")))

json.dump({
    "source": "examples/README.md:7-12; examples/multivariate_normal/mvn_example.ipynb cell 4; mvn_example.cpp:9-39",
    "dim": 2, "num_particles": 10, "num_iterations": 1000, "x0_scale": 3.0,
    "x0": "scale * Eigen::MatrixXd::Random(dim, n), unseeded glibc rand()",
    "means": [[-0.6871, 0.8010]],
    "covs": [[[5 * 0.2260, 5 * 0.1652], [5 * 0.1652, 5 * 0.6779]]],
    "optimizer": {"kind": "adagrad", "lr": 0.1, "eps": 1e-8},
    "kernel": "GaussianRBFKernel, ScaleMethod::Median",
    "printed_significant_digits": 6,
    "initial": mvn_init, "final": mvn_final,
}, open(os.path.join(HERE, "mvn_example.json"), "w"), indent=1)

json.dump({
    "source": "examples/gaussian_mixture_model/gmm_example.ipynb cell 4; gmm_example.cpp:9-49",
    "dim": 2, "num_particles": 20, "num_iterations": 1000, "x0_scale": 8.0,
    "x0": "scale * Eigen::MatrixXd::Random(dim, n), unseeded glibc rand()",
    "means": [[3.6871, -2.801], [-2.9802, 4.3387]],
    "covs": [[[5 * 0.5001, 5 * 0.2426], [5 * 0.2426, 5 * 0.8420]],
             [[5 * 0.6779, 5 * -0.1652], [5 * -0.1652, 5 * 0.2260]]],
    "optimizer": {"kind": "adam", "lr": 0.1, "beta1": 0.9, "beta2": 0.999, "eps": 1e-8},
    "kernel": "GaussianRBFKernel, ScaleMethod::Median",
    "printed_significant_digits": 6,
    "initial": gmm_init, "final": gmm_final,
}, open(os.path.join(HERE, "gmm_example.json"), "w"), indent=1)
print("wrote mvn_example.json, gmm_example.json")
