#!/usr/bin/env python3
"""Extract the two recorded 2-body traces of the reference's manual accuracy study.

Source: /root/reference/tests/accuracy.ods (sheets ``data-600`` and ``data-60000``; see
SURVEY.md section 4).  The sheets hold the stderr trace of an old debug build, one row per
body per step: ``time;name;pos;vel`` with vectors printed as ``(x, y, z, 1)`` at 6 significant
figures.  These are the only golden vectors the reference ships for the World::update path.

Output (committed test data, inputs only -- no reference code):
    tests/golden/accuracy_data-600.csv.gz
    tests/golden/accuracy_data-60000.csv.gz
Columns: time,name,px,py,pz,vx,vy,vz  (numbers kept as the literal strings of the sheet).

Run in the authoring container only (the GPU box has no /root/reference):
    python tools/extract_accuracy_ods.py
"""
import gzip
import os
import re
import sys
import zipfile

SRC = "/root/reference/tests/accuracy.ods"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

VEC = re.compile(r"\(([^,]+), ([^,]+), ([^,]+), 1\)")


def rows_of(sheet_xml):
    for row in re.finditer(r"<table:table-row[^>]*>(.*?)</table:table-row>", sheet_xml, re.S):
        cells = re.findall(r"<text:p>(.*?)</text:p>", row.group(1), re.S)
        if len(cells) < 4 or not cells[0].isdigit():
            continue
        p, v = VEC.fullmatch(cells[2]), VEC.fullmatch(cells[3])
        if not (p and v):
            continue
        yield [cells[0], cells[1], *p.groups(), *v.groups()]


def main():
    content = zipfile.ZipFile(SRC).read("content.xml").decode("utf-8")
    for sheet in ("data-600", "data-60000"):
        m = re.search(r'<table:table table:name="%s".*?</table:table>' % re.escape(sheet), content, re.S)
        if not m:
            sys.exit("sheet %s not found" % sheet)
        rows = list(rows_of(m.group(0)))
        path = os.path.join(OUT, "accuracy_%s.csv.gz" % sheet)
        with gzip.GzipFile(path, "wb", mtime=0) as raw:
            raw.write(b"time,name,px,py,pz,vx,vy,vz\n")
            for r in rows:
                raw.write((",".join(r) + "\n").encode())
        print(sheet, len(rows), "rows ->", os.path.normpath(path))


if __name__ == "__main__":
    main()
