"""Hand-written golden sites of SURVEY.md section 8c: (ref_char, [(base_char, phred, strand_char), ...]).

G1..G7 exercise the ordinary paths (bi-allelic, mixed quals, mono-allelic 5000 rule, tetra-allelic, min_af pruning),
E1..E6 the edge cases (phred 0 -> NaN AF, indels only, lowercase / 'N' reference, chi2 survival underflow -> QUAL 10000).
The reference's outputs for them live in sites_minaf_*.npz (made by make_golden.py from the compiled reference); the
literal values quoted in SURVEY.md 8c (probe of the reference binary) are asserted in tests/test_oracle_golden.py.
"""
GOLDEN_MIN_AF = [0.01, 0.05, 0.001]

GOLDEN_SITES = [
    # G1
    ("A", [("A", 30, "-+"[i % 2]) for i in range(7)] + [("G", 30, "-+"[i % 2]) for i in range(3)]),
    # G2
    ("C", [("C", 20 + i % 20, "+" if i % 3 else "-") for i in range(50)] + [("T", 35, "+")] * 5 + [("A", 12, "-")] * 2
     + [("N", 0, ".")] * 20),
    # G3
    ("G", [("T", 30, "+")] * 12),
    # G5
    ("A", [("A", 30, "-+"[i % 2]) for i in range(40)] + [("C", 30, "-+"[i % 2]) for i in range(30)] + [("G", 30, "+")] * 20
     + [("T", 30, "-")] * 10),
    # G6/G7
    ("A", [("A", 30 + i % 10, "+-"[i % 2]) for i in range(997)] + [("C", 25, "+")] * 3),
    # E1
    ("G", [("T", 0, "+")]),
    # E2
    ("A", [("A", 30, "+")] * 5 + [("G", 0, "+")] * 3),
    # E3
    ("A", [("+", 30, "+"), ("-", 30, "+"), ("N", 0, ".")]),
    # E4
    ("a", [("A", 30, "+")] * 2 + [("C", 30, "+")] * 2 + [("C", 35, "+")]),
    # E5
    ("N", [("A", 30, "+")] * 2 + [("C", 30, "+")] * 2 + [("C", 35, "+")]),
    # E6
    ("A", [("A", 40, "+")] * 600 + [("T", 40, "-")] * 400),
    # empty site
    ("T", []),
    # G4 (range.bam VCF rows)
    ("G", [("T", 37, "-")] * 2),
]

# Literal outputs of the reference binary quoted in SURVEY.md 8c (as-built / -include stdlib.h), %.17g.
# index into GOLDEN_SITES -> (min_af, alts, af_int, qual_int, af_dbl, qual_dbl)
SURVEY_LITERALS = {
    0: (0.01, "G", [0.2998664889880594], 86.650670492227547, [0.29986648865206966], 86.650670492232265),
    1: (0.01, "T", [0.090898670649150692], 135.99914625292956, [0.090898656745670575], 135.99914625315017),
    2: (0.01, "T", [1.0], 5000.0, [1.0], 5000.0),
    3: (0.01, "CGT", [0.30006675422865065, 0.1999332441682892, 0.099799736856721194], 217.57830376056739,
        [0.30006675566795432, 0.1999332443234654, 0.099799732994476267], 217.57830359121249),
    4: (0.01, "", [], 0.0, [], 0.0),
    6: (0.01, "", [], 0.0, [], 0.0),
    7: (0.01, "", [], 0.0, [], 0.0),
    8: (0.01, "C", [0.60009720652596177], 63.079116168710613, None, None),
    9: (0.01, "AC", [0.39990279347403834, 0.60009720652596177], 63.079116168710613, None, None),
    10: (0.01, "T", [0.39999333244445429], 10000.0, None, None),
    12: (0.05, "T", [1.0], 0.0, [1.0], 0.0),
}
