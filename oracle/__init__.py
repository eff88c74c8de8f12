"""CPU oracle (test infrastructure only).  See oracle/render_oracle.py for the import rules."""
