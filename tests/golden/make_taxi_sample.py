"""Extracts the five columns taxi Q1-Q4 read from the reference's 20-row taxi sample
(omniscidb/Tests/ArrowStorageDataFiles/taxi_sample.csv, imported without a header by ArrowStorageTaxiTest,
omniscidb/Tests/ArrowStorageSqlTest.cpp:131-191) into tests/golden/taxi_sample_columns.json.
Run in the build container (it reads /root/reference); the JSON is what travels to the GPU box."""
import csv
import json
import os

SRC = "/root/reference/omniscidb/Tests/ArrowStorageDataFiles/taxi_sample.csv"
# 0-based positions in the schema of ArrowStorageSqlTest.cpp:139-189
COLS = {"pickup_datetime": 2, "passenger_count": 10, "trip_distance": 11, "total_amount": 19, "cab_type": 24}

if __name__ == "__main__":
    rows = list(csv.reader(open(SRC)))
    out = {name: [r[i] for r in rows] for name, i in COLS.items()}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "taxi_sample_columns.json")
    json.dump(out, open(dst, "w"), indent=0)
    print(f"{len(rows)} rows -> {dst}")
