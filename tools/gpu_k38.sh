#!/bin/bash
echo "## 3 CTAs per SM (80 registers)"; python tools/time_basis_combine.py 2>/dev/null
echo "## 2 CTAs per SM (98 registers)"; DMH_LIB=tools/_var/bc2/libdmhomo.so python tools/time_basis_combine.py 2>/dev/null
