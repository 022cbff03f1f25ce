#!/bin/bash
bash tools/gpu_refnets.sh
