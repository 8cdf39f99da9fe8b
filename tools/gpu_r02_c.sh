#!/bin/bash
timeout 600 python tools/profile_suggest.py 2>&1 | tail -45
