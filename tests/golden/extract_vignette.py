"""Extract the known answers the reference itself published for the hot path (run in the BUILD container,
where /root/reference exists):

    python tests/golden/extract_vignette.py

Source: the reference's built vignette doc/plaid-vignette.html (from vignettes/plaid-vignette.Rmd:89-168), which
prints, for the bundled fixture (inst/extdata/pbmc3k-50cells.rda x hallmarks.gmt):
  * dim(matG) and dim(gsetX)                                                        (Rmd :89-109)
  * head(res) of  res <- plaid.test(X, y, matG, gsetX = gsetX, tests = c("one", "lm"))    (Rmd :150-160)
    with  gsetX <- plaid(X, matG, normalize = TRUE)  and  y <- 1 * (celltype == "B").
The printed columns `p.lm` (Welch two-group t-test on the rows of gsetX, Rfast::ttests, R/plaid.R:429) and `p.one`
(one-sample t-test on the per-gene log fold changes inside each set, R/plaid.R:476-520) are reproduced by the
current code's arithmetic to all 7 printed digits; `gsetFC`, `p.meta` and `q.meta` were printed by an older
revision of plaid.test (different fold-change summary and p-value combination) and are recorded but not used.
Output: tests/golden/vignette_known_answers.json.
"""
import html
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PLAID_REFERENCE", "/root/reference")


def main():
    s = open(os.path.join(REF, "doc", "plaid-vignette.html"), encoding="utf-8").read()
    s = re.sub(r"<script.*?</script>|<style.*?</style>", "", s, flags=re.S)
    t = html.unescape(re.sub(r"<[^>]+>", "", s))
    dims = re.findall(r"#> \[1\]\s+(\d+)\s+(\d+)", t)
    rows = {}
    # head(res) is printed in two blocks (3 + 2 columns), each line "#> NAME v1 v2 ..."
    for m in re.finditer(r"#> (HALLMARK_\w+)\s+([-0-9.e+ ]+)", t):
        rows.setdefault(m.group(1), []).extend(float(v) for v in m.group(2).split())
    cols = ["gsetFC", "p.one", "p.lm", "p.meta", "q.meta"]
    table = {k: dict(zip(cols, v)) for k, v in rows.items() if len(v) == len(cols)}
    out = {
        "source": "bigomics/plaid doc/plaid-vignette.html (built from vignettes/plaid-vignette.Rmd:89-168)",
        "call": 'gsetX <- plaid(X, matG, normalize=TRUE); y <- 1*(celltype == "B"); '
                'res <- plaid.test(X, y, matG, gsetX=gsetX, tests=c("one","lm")); head(res[order(res[,"p.meta"]),])',
        "dims_printed": [[int(a), int(b)] for a, b in dims],
        "head_res": table,
        "pinned_columns": ["p.one", "p.lm"],
        "stale_columns": ["gsetFC", "p.meta", "q.meta"],
    }
    path = os.path.join(HERE, "vignette_known_answers.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path, "-", len(table), "rows; dims", out["dims_printed"])


if __name__ == "__main__":
    main()
