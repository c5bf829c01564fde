#!/bin/bash
timeout 120 tools/micro/bin/potrf_check 2>&1 | grep -E "blocked kernel|last diagonal|variant 2: "
