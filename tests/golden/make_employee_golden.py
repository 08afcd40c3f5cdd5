"""Derive golden vectors from the reference's only fixture, kquerydiy/employee.csv.

Run in the build container (needs /root/reference):  python tests/golden/make_employee_golden.py
The reference ships no tests or expected outputs; these vectors are what its code
yields on that file, derived by hand from the rules in SURVEY.md §8c — NOT with the oracle:
  - CsvDataSource: header row -> all-Utf8 fields; values trimmed; missing -> "" (Main.kt:263, 345-348)
  - SELECT state, MAX(CAST(salary AS double)) GROUP BY state (the shape of Main.kt:1336)
  - BASELINE config 1: SELECT id, first_name, last_name, state, salary WHERE state = 'CO'
"""
import csv
import json
import os

SRC = "/root/reference/kquerydiy/employee.csv"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "employee_golden.json")

with open(SRC, newline="", encoding="utf-8") as f:
    rows = list(csv.reader(f))
header = [h.strip() for h in rows[0]]
data = [[c.strip() for c in r] for r in rows[1:] if r]
cols = {h: [r[i] if i < len(r) else "" for r in data] for i, h in enumerate(header)}

# GROUP BY state, MAX(CAST(salary AS double)): first non-null initialises, strict '>' replaces (Main.kt:540-555)
mx = {}
for st, sal in zip(cols["state"], cols["salary"]):
    v = float(sal)
    if st not in mx or v > mx[st]:
        mx[st] = v

def where_state(val):
    keep = [i for i, s in enumerate(cols["state"]) if s == val]
    return {k: [cols[k][i] for i in keep] for k in ("id", "first_name", "last_name", "state", "salary")}

golden = {
    "source": "kquerydiy/employee.csv (157 bytes) read with csv.reader + strip()",
    # the fixture's bytes (data, not code), so that the CSV scan tests can run where /root/reference does not exist
    "csv_text_hex": open(SRC, "rb").read().hex(),
    "schema": header,
    "columns": cols,
    "last_name_row3_utf8_hex": cols["last_name"][2].encode("utf-8").hex(),
    "group_by_state_max_salary": mx,
    "config1_where_state_eq_CO": where_state("CO"),
    "where_state_eq_Uppsala": where_state("Uppsala"),
}
with open(OUT, "w", encoding="utf-8") as f:
    json.dump(golden, f, indent=1, ensure_ascii=False, sort_keys=True)
print("wrote", OUT)
