"""The BASELINE.json workloads written against this package's host API (each mirrors the set-up of the
reference example of the same name and produces the same encoded geometry)."""
