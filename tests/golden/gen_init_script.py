#!/usr/bin/env python
"""Extracts the wait-gcs-ready polling script from the reference source (a Go raw string inside DefaultWorkerPodTemplate,
ray-operator/controllers/ray/common/pod.go) and writes it, byte for byte with its %s placeholders, to wait_gcs_ready_script.json.
Run in the build container (the reference tree does not travel to the GPU box); the JSON is committed."""
import hashlib
import json
import os
import re

SRC = "/root/reference/ray-operator/controllers/ray/common/pod.go"
HERE = os.path.dirname(os.path.abspath(__file__))

text = open(SRC, encoding="utf-8").read()
start = text.index('Name:            "wait-gcs-ready"')
m = re.compile(r"fmt\.Sprintf\(`(.*?)`, fqdnRayIP, headPort, fqdnRayIP, headPort\)", re.S).search(text, start)
assert m, "the init container's script literal was not found"
script = m.group(1)
line0 = text.count("\n", 0, m.start(1)) + 1
doc = {"source": f"common/pod.go:{line0}-{line0 + script.count(chr(10))}", "placeholders": "fqdnRayIP, headPort, fqdnRayIP, headPort",
       "sha256": hashlib.sha256(script.encode()).hexdigest(), "format": script}
with open(os.path.join(HERE, "wait_gcs_ready_script.json"), "w") as f:
    json.dump(doc, f, indent=1)
    f.write("\n")
print(doc["source"], doc["sha256"], len(script), "bytes")
