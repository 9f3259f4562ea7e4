#=GENOME_DIFF	1.0
RA	1	.	edge	1	0	A	.	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=.	major_cov=28/30	minor_base=A	minor_cov=1/1	new_cov=28/30	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
RA	2	.	edge	4	1	.	A	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=A	major_cov=28/30	minor_base=.	minor_cov=1/1	new_cov=28/30	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
RA	5	.	edge	9	0	G	.	consensus_reject=FREQUENCY_CUTOFF	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	ks_quality_p_value=5.00000e-01	major_base=G	major_cov=20/22	minor_base=.	minor_cov=9/9	new_cov=9/9	prediction=polymorphism	ref_cov=20/22	score=40.0	total_cov=30/30
RA	9	.	edge	22	0	C	T	consensus_reject=EXISTING,FREQUENCY_CUTOFF	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	ks_quality_p_value=5.00000e-01	major_base=C	major_cov=20/22	minor_base=T	minor_cov=9/9	new_cov=9/9	prediction=polymorphism	ref_cov=20/22	score=40.0	total_cov=30/30	user_defined=1
RA	14	.	edge	28	1	.	A	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=A	major_cov=28/30	minor_base=.	minor_cov=1/1	new_cov=28/30	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
RA	15	.	edge	34	0	A	T	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=T	major_cov=28/30	minor_base=A	minor_cov=1/1	new_cov=28/30	note=a=b	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
RA	16	.	edge	39	1	.	C	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=C	major_cov=28/30	minor_base=.	minor_cov=1/1	new_cov=28/30	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
RA	17	.	edge	39	2	.	G	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=G	major_cov=28/30	minor_base=.	minor_cov=1/1	new_cov=28/30	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
MC	18	.	edge	1	2	0	0	left_inside_cov=0	left_outside_cov=NA	right_inside_cov=0	right_outside_cov=5
UN	19	.	edge	1	2
