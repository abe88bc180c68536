"""Extracts the pairs() figure of the reference's built vignette (doc/plaid-vignette.html, section 4.4 "Compare
scores": `S <- cbind(plaid=gsetX[,1], sing=sing[,1], ssgsea=ssgsea[,1], scSE=scse[,1]); pairs(S)`, from
vignettes/plaid-vignette.Rmd:107,202,217,235,252) into tests/golden/vignette_pairs.png.  The figure is reference OUTPUT:
the only place where the reference published results of replaid.sing / replaid.ssgsea / replaid.scse (and, through
them, of colranks / sparse_colranks) for its bundled fixture.  Run in the build container (needs /root/reference)."""
import base64
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
html = open("/root/reference/doc/plaid-vignette.html").read()
imgs = re.findall(r'<img src="data:image/png;base64,([A-Za-z0-9+/=]+)"', html)
assert len(imgs) == 2, len(imgs)  # [0] volcano plot of plaid.test, [1] the pairs plot
open(os.path.join(HERE, "vignette_pairs.png"), "wb").write(base64.b64decode(imgs[1]))
print("wrote vignette_pairs.png", len(base64.b64decode(imgs[1])), "bytes")
