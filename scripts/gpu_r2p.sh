#!/bin/bash
bash scripts/profile_r2.sh r2p
