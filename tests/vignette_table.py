"""The top-10 table the reference prints for `set.seed(1234); haystack(dat.tsne, dat.expression)`
(/root/reference/docs/articles/a01_toy_example.html:168), copied from the rendered page."""

# docs/articles/a01_toy_example.html:168 (gene, D_KL, log.p.vals, log.p.adj), in printed order
VIGNETTE_TOP10 = [
    ("gene_79", 2.447641, -39.94871, -37.24974),
    ("gene_497", 2.271242, -39.84546, -37.14649),
    ("gene_242", 1.742783, -35.91184, -33.21287),
    ("gene_275", 1.819669, -35.88741, -33.18844),
    ("gene_62", 2.174074, -35.63854, -32.93957),
    ("gene_71", 2.546493, -34.83522, -32.13625),
    ("gene_381", 2.733446, -34.38544, -31.68647),
    ("gene_351", 1.844673, -33.45724, -30.75827),
    ("gene_479", 2.343509, -31.98940, -29.29043),
    ("gene_300", 2.097546, -30.42637, -27.72740),
]
