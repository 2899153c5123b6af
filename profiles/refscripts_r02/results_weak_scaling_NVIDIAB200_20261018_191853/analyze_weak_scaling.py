#!/usr/bin/env python3
"""
Analyze weak scaling results (constant work per GPU)
"""
import json
import glob
import matplotlib.pyplot as plt
import numpy as np

# Read all JSON files
results = []
for json_file in sorted(glob.glob("*.json")):
    with open(json_file) as f:
        # Read and fix inf values
        content = f.read().replace(': inf', ': null')
        data = json.loads(content)

        # Use grid_size (not rows)
        grid_size = data['matrix']['grid_size']
        results.append({
            'gpus': data['num_gpus'],
            'size': grid_size,
            'unknowns': data['matrix']['rows'],
            'unknowns_per_gpu': data['matrix']['rows'] // data['num_gpus'],
            'iterations': data['convergence']['iterations'],
            'time_ms': data['timing']['median_ms'],
            'time_per_iter': data['timing']['median_ms'] / data['convergence']['iterations']
        })

# Sort by GPU count
results = sorted(results, key=lambda x: x['gpus'])

# Create weak scaling plots (2-panel)
fig, (ax1, ax2) = plt.subplots(1, 2, figsize=(14, 5))

colors = '#2E86AB'
gpus_list = [r['gpus'] for r in results]
times = [r['time_ms'] for r in results]
sizes = [r['size'] for r in results]

# Plot 1: Time vs GPU count (should be ~constant for perfect weak scaling)
baseline_time = results[0]['time_ms']
ax1.plot(gpus_list, times, marker='o', linestyle='-',
         color=colors, linewidth=2.5, markersize=10,
         markeredgewidth=1.5, markeredgecolor='white', label='Measured')
ax1.axhline(y=baseline_time, color='k', linestyle='--', linewidth=2,
            alpha=0.6, label=f'Ideal ({baseline_time:.1f} ms)')

ax1.set_xlabel('Number of GPUs', fontsize=13, fontweight='bold')
ax1.set_ylabel('Total Time (ms)', fontsize=13, fontweight='bold')
ax1.set_title('Weak Scaling: Time vs GPU Count', fontsize=14, fontweight='bold', pad=15)
ax1.set_xticks(gpus_list)
ax1.legend(fontsize=11, frameon=True, shadow=True)
ax1.grid(True, alpha=0.3, linestyle='--')
ax1.set_ylim([0, max(times) * 1.2])

# Plot 2: Weak scaling efficiency
efficiencies = [100 * baseline_time / t for t in times]
ax2.bar(range(len(gpus_list)), efficiencies, color=colors, alpha=0.8,
        edgecolor='black', linewidth=1.5)
ax2.axhline(y=100, color='k', linestyle='--', linewidth=2, alpha=0.6, label='Ideal')

ax2.set_xticks(range(len(gpus_list)))
ax2.set_xticklabels([f'{g} GPU{"s" if g>1 else ""}' for g in gpus_list])
ax2.set_ylabel('Weak Scaling Efficiency (%)', fontsize=13, fontweight='bold')
ax2.set_title('Weak Scaling Efficiency', fontsize=14, fontweight='bold', pad=15)
ax2.legend(fontsize=11, frameon=True, shadow=True)
ax2.grid(True, alpha=0.3, axis='y', linestyle='--')
ax2.set_ylim([0, 110])

# Add value labels on bars
for i, (eff, time) in enumerate(zip(efficiencies, times)):
    ax2.text(i, eff + 2, f'{eff:.1f}%\n({time:.1f}ms)',
             ha='center', va='bottom', fontsize=10, fontweight='bold')

plt.tight_layout()
plt.savefig('weak_scaling_analysis.png', dpi=300, bbox_inches='tight', facecolor='white')
print("✅ Generated: weak_scaling_analysis.png")

# Print summary table
print("\n" + "="*100)
print("Weak Scaling Results (Constant ~25M unknowns per GPU)")
print("="*100)
print(f"{'GPUs':<6} {'Size':<10} {'Total Unknowns':<15} {'Per GPU':<12} {'Iter':<6} {'Time (ms)':<12} {'Efficiency':<12}")
print("-"*100)

for r in results:
    efficiency = 100 * baseline_time / r['time_ms']
    print(f"{r['gpus']:<6} {r['size']:<10} {r['unknowns']:<15,} {r['unknowns_per_gpu']:<12,} "
          f"{r['iterations']:<6} {r['time_ms']:<12.2f} {efficiency:<12.1f}%")

print("-"*100)
print(f"\nBaseline (1 GPU): {baseline_time:.2f} ms")
print(f"Ideal weak scaling: constant time ({baseline_time:.2f} ms) for all GPU counts")
print("="*100)
