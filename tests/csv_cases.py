"""CSV texts for the scan tests (CsvDataSource, Main.kt:276-357; rules C1-C9 in csrc/kq_csv.cu) and a generator of
large synthetic files. Expected values of the hand-written cases are spelled out so that the oracle itself is pinned."""
import numpy as np

# (name, text, has_headers, expected header names, expected columns as lists of str)
CASES = [
    ("plain", b"a,b,c\n1,2,3\n4,5,6\n", True, ["a", "b", "c"], [["1", "4"], ["2", "5"], ["3", "6"]]),
    ("no_trailing_newline", b"a,b\n1,2\n3,4", True, ["a", "b"], [["1", "3"], ["2", "4"]]),
    ("crlf", b"a,b\r\n1,2\r\n3,4\r\n", True, ["a", "b"], [["1", "3"], ["2", "4"]]),
    ("cr_only", b"a,b\r1,2\r3,4\r", True, ["a", "b"], [["1", "3"], ["2", "4"]]),
    ("semicolon", b"a;b;c\r\n1; \"x;y\" ;3\r\n\r\n\"q\"\"r\";;\r\n5;6", True, ["a", "b", "c"], [["1", 'q"r', "5"], ["x;y", "", "6"], ["3", "", ""]]),
    ("tab", b"x\ty\n 1 \t 2 \n", True, ["x", "y"], [["1"], ["2"]]),
    ("pipe", b"x|y|z\n1|2|3\n", True, ["x", "y", "z"], [["1"], ["2"], ["3"]]),
    ("trim", b"a,b\n  padded  ,\t tab \n", True, ["a", "b"], [["padded"], ["tab"]]),
    ("quoted_newline", b'a,b\n"line1\nline2",2\n3,"x,y"\n', True, ["a", "b"], [["line1\nline2", "3"], ["2", "x,y"]]),
    ("quoted_trim", b'a\n"  spaced  "\n" ""q"" "\n', True, ["a"], [["spaced", '"q"']]),
    ("after_quote_dropped", b'a,b\n"v"junk,2\n', True, ["a", "b"], [["v"], ["2"]]),
    ("empty_lines", b"\n\na,b\n\n1,2\n\n\n3,4\n\n", True, ["a", "b"], [["1", "3"], ["2", "4"]]),
    ("short_and_long_rows", b"a,b,c\n1\n1,2,3,4,5\n,,\n", True, ["a", "b", "c"], [["1", "1", ""], ["", "2", ""], ["", "3", ""]]),
    ("no_headers", b"1,2\n3,4\n", False, ["field_1", "field_2"], [["1", "3"], ["2", "4"]]),
    ("header_only", b"a,b\n", True, ["a", "b"], [[], []]),
    ("utf8", "namn,ort\nPärsson,Åre\n日本,東京\n".encode("utf-8"), True, ["namn", "ort"], [["Pärsson", "日本"], ["Åre", "東京"]]),
    ("single_column", b"v\n1\n2\n3\n", True, ["v"], [["1", "2", "3"]]),
    ("blank_field_line", b"a\n \n1\n", True, ["a"], [["", "1"]]),
    # rule C2: a quote opens a quoted section only as the first non-blank byte of a field; anywhere else it is data
    ("stray_quote_is_data", b'a,b\n5" pipe,x\n6,y"z\n', True, ["a", "b"], [['5" pipe', "6"], ["x", 'y"z']]),
    ("stray_quote_keeps_lines", b'a,b\nit"s,1\nnext,2\n"q,""r""\n",3\n', True, ["a", "b"], [['it"s', "next", 'q,"r"'], ["1", "2", "3"]]),
    ("quote_after_closed_section", b'a,b\n"v" "w,2\n"x""",3\n', True, ["a", "b"], [["v", 'x"'], ["2", "3"]]),
    ("quoted_after_blanks", b'a;b\n1;  "x;y"  \n\t"p""";2\n', True, ["a", "b"], [["1", 'p"'], ["x;y", "2"]]),
    ("stray_quote_in_header", b'wi"dth,he"ight\n1,2\n', True, ['wi"dth', 'he"ight'], [["1"], ["2"]]),
]


def synthetic(n_rows, seed=7, ncols=6, quoted_every=13, crlf=False):
    """A CSV file of n_rows records: ids, short words, numbers with padding, and a quoted field with delimiters,
    doubled quotes and (rarely) a line break inside. Returns bytes."""
    rng = np.random.default_rng(seed)
    words = ["Uppsala", "Sthlm", "CO", "Åre", "x", "", "Eng", "Worker"]
    eol = "\r\n" if crlf else "\n"
    out = [",".join(f"c{i}" for i in range(ncols)) + eol]
    w = rng.integers(0, len(words), size=(n_rows, 2))
    v = rng.integers(0, 10**9, size=n_rows)
    pad = rng.integers(0, 3, size=n_rows)
    for r in range(n_rows):
        f = [str(r), words[w[r, 0]], " " * pad[r] + str(v[r]) + " " * pad[r], words[w[r, 1]]]
        if r % quoted_every == 0:
            f.append('"a,b ""%d""%s"' % (r, "\n" if r % (quoted_every * 7) == 0 else ""))
        else:
            f.append("plain%d" % (r % 97))
        f += [str((r * 31) % 1000)] * (ncols - len(f))
        out.append(",".join(f[:ncols]) + eol)
        if r % 1001 == 1000:
            out.append(eol)            # an empty line now and then
    return "".join(out).encode("utf-8")
