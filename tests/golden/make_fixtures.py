#!/usr/bin/env python
"""Regenerates tests/golden/*.gz from the reference's bundled dataset.

The golden vectors are DATA (not source) shipped in /root/reference/dataset
(dataset/README.md:12-17); the reference's own tests compare app output with
them (misc/app_tests.sh:51-113).  They are committed gzip-compressed because
/root/reference does not exist on the GPU box.
"""
import gzip, os, shutil, sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/dataset"
DST = os.path.dirname(os.path.abspath(__file__))
FILES = ["p2p-31.e", "p2p-31.v", "p2p-31-BFS", "p2p-31-BFS-directed",
         "p2p-31-SSSP", "p2p-31-SSSP-directed", "p2p-31-PR",
         "p2p-31-PR-directed", "p2p-31-CDLP", "p2p-31-LCC", "p2p-31-WCC"]
for f in FILES:
    with open(os.path.join(SRC, f), "rb") as i, \
         gzip.GzipFile(os.path.join(DST, f + ".gz"), "wb", 9, mtime=0) as o:
        shutil.copyfileobj(i, o)
    print("wrote", f + ".gz")
