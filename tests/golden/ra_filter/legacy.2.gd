#=GENOME_DIFF	1.0
RA	3	.	edge	5	0	T	A	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=A	major_cov=12/14	minor_base=T	minor_cov=1/1	new_cov=28/30	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
RA	9	.	edge	22	0	C	T	consensus_reject=EXISTING,FREQUENCY_CUTOFF	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	ks_quality_p_value=5.00000e-01	major_base=C	major_cov=20/22	minor_base=T	minor_cov=9/9	new_cov=9/9	prediction=polymorphism	ref_cov=20/22	score=40.0	total_cov=30/30	user_defined=1
RA	10	.	edge	23	0	G	A	consensus_reject=FREQUENCY_CUTOFF	consensus_score=5.0	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	ks_quality_p_value=5.00000e-01	major_base=G	major_cov=20/22	minor_base=A	minor_cov=9/9	new_cov=9/9	polymorphism_score=30.0	prediction=polymorphism	ref_cov=20/22	total_cov=10/12
RA	11	.	edge	24	0	C	A	consensus_reject=FREQUENCY_CUTOFF	fisher_strand_p_value=5.00000e-01	frequency=3.0e-01	ks_quality_p_value=5.00000e-01	major_base=C	major_cov=20/22	minor_base=A	minor_cov=3/3	new_cov=9/9	prediction=polymorphism	ref_cov=20/22	score=NA	total_cov=30/30
RA	14	.	edge	28	1	.	A	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=A	major_cov=28/30	minor_base=.	minor_cov=1/1	new_cov=28/30	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
RA	16	.	edge	39	1	.	C	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=C	major_cov=28/30	minor_base=.	minor_cov=1/1	new_cov=28/30	prediction=consensus	ref_cov=1/1	score=40.0	total_cov=30/30
RA	17	.	edge	39	2	.	G	consensus_reject=INDEL_HOMOPOLYMER	fisher_strand_p_value=5.00000e-01	frequency=1.0e+00	ks_quality_p_value=5.00000e-01	major_base=G	major_cov=28/30	minor_base=.	minor_cov=1/1	new_cov=28/30	prediction=polymorphism	ref_cov=1/1	score=40.0	total_cov=30/30
MC	18	.	edge	1	2	0	0	left_inside_cov=0	left_outside_cov=NA	right_inside_cov=0	right_outside_cov=5
UN	19	.	edge	1	2
