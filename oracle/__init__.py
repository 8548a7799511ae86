"""TEST INFRASTRUCTURE ONLY: CPU restatements of the reference algorithm (the parity oracle); never imported by the product."""
