"""hoh-ANS hot path on B200: host-side mirror of the reference's function interface over libhohgpu.so."""
