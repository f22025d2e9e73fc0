#!/usr/bin/env python
"""Generates tests/golden/result_files.json by EXECUTING the reference's own result-file helpers
(test.py: load_txt_to_dict :1650-1658, save_dict_to_txt :1660-1664, update_txt_file :1666-1674, process_line
:1788-1796, and the clean-up loop of run_test1 :1843-1849) on crafted result files.  TEST INFRASTRUCTURE ONLY; run in the
build container (it reads /root/reference).  No reference source is stored: only input / output strings.

    python oracle/make_golden_results.py [--reference /root/reference]
"""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.make_golden import extract  # noqa: E402

CASES = {
    # base list (evaluate_base, test.py:1742-1747) overwritten per file by the OOD list (evaluate_new, :1776-1781)
    "override_and_append": (
        "['TestSetB/a.jpg'] 1 2 3 4 5\n['TestSetB/b.jpg'] 6 7 8 9 10\n['TestSetB/c.jpg'] 11 12 13 14 15\n",
        "['TestSetB/b.jpg'] 374 375 376 377 378\n['TestSetB/z.jpg'] 400 401 402 399 398\n"),
    "empty_update": ("['x/1.jpg'] 0 1 2 3 4\n", ""),
    "empty_base": ("", "['x/1.jpg'] 0 1 2 3 4\n['x/2.jpg'] 5 4 3 2 1\n"),
    "duplicate_keys_last_wins": ("['d/q.jpg'] 1 1 1 1 1\n['d/q.jpg'] 2 2 2 2 2\n['d/r.jpg'] 3 3 3 3 3\n",
                                 "['d/r.jpg'] 9 9 9 9 9\n['d/r.jpg'] 8 8 8 8 8\n"),
    "irregular_whitespace_and_short_rows": ("['p/a.png']   1  2 3\t4 5  \n['p/b.png'] 7\n", "['p/c.png'] 1 2 3 4 5\n"),
    "nested_dirs_and_plain_names": ("['Dataset/TestSetB/sub/dir/img_001.jpg'] 1 2 3 4 5\nplain.jpg 5 4 3 2 1\n",
                                    "['Dataset/TestSetB/sub/dir/img_001.jpg'] 10 20 30 40 50\n"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "result_files.json"))
    a = ap.parse_args()
    ns = {}
    exec("import re\n" + extract(os.path.join(a.reference, "test.py"),
                                 ["load_txt_to_dict", "save_dict_to_txt", "update_txt_file", "process_line"]), ns)
    out = {}
    for name, (base, update) in CASES.items():
        with tempfile.TemporaryDirectory() as d:
            b, u, r = (os.path.join(d, n) for n in ("base.txt", "update.txt", "result.txt"))
            open(b, "w").write(base)
            open(u, "w").write(update)
            ns["update_txt_file"](b, u)                                   # test.py:1840
            merged = open(b).read()
            with open(b) as fi, open(r, "w") as fo:                       # test.py:1846-1849
                for line in fi:
                    fo.write(ns["process_line"](line))
            out[name] = {"base": base, "update": update, "merged": merged, "result": open(r).read()}
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {a.out}: {len(out)} cases")


if __name__ == "__main__":
    main()
