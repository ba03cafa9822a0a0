"""oracle/ -- TEST INFRASTRUCTURE ONLY (checker for the CUDA path; never shipped, never timed
as the product).  See oracle/README.md."""
