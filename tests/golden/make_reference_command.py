"""Extracts the container `command:` and `image:` of the reference Deployment
(/root/reference/cuda-test-deployment.yaml:18-19) into tests/golden/reference_command.json.

The reference tree does not travel to the GPU box, so the parsed values are committed as a
fixture; tests/test_image_layout.py re-parses the YAML whenever /root/reference is present and
fails if the fixture has drifted.

    python tests/golden/make_reference_command.py
"""
import json
import os

import yaml

REF = "/root/reference/cuda-test-deployment.yaml"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_command.json")


def parse(path: str = REF) -> dict:
    doc = yaml.safe_load(open(path))
    c = doc["spec"]["template"]["spec"]["containers"][0]
    return {"source": "cuda-test-deployment.yaml:18-19", "container": c["name"], "image": c["image"], "command": c["command"],
            "gpu_limit": c["resources"]["limits"]["nvidia.com/gpu"]}


if __name__ == "__main__":
    json.dump(parse(), open(OUT, "w"), indent=1)
    print(open(OUT).read())
